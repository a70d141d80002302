"""Mirror of the reference's `RTAB_utils` namespace (only what the label-fusion path touches)."""
