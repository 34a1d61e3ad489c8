"""Host-side mirror of the reference's `lib/` package, restricted to the render hot path."""
