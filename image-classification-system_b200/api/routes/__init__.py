"""Host-side mirrors of the reference's route handlers for the hot path (app/api/routes)."""
