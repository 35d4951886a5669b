"""Host-side mirrors of the reference's utils package (transform chains, Jacobians, trajectory generators,
calibration) with batched GPU entry points; see INTEGRATION.md for the call table."""
