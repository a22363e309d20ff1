"""Adapters between the hot path's persistence seam (``store.ImageStore``) and the reference's storage layer."""
