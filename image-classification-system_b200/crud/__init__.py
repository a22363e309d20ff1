"""Host-side mirrors of the reference's CRUD functions for the hot path (app/crud)."""
