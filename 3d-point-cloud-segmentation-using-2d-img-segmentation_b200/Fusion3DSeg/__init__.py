"""Mirror of the reference's `Fusion3DSeg` namespace (same module and function names)."""
