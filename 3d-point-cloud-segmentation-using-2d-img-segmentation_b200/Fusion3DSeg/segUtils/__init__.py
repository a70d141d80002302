"""Mirror of the reference's `Fusion3DSeg.segUtils` namespace."""
