"""Host-side mirrors of the reference's service classes for the hot path (app/services)."""
