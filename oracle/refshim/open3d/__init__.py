"""TEST INFRASTRUCTURE ONLY -- empty stand-in so that `import open3d as o3d` at the top of the reference's
`Fusion3DSeg/fusion.py:6` succeeds in the build container (Open3D is not installed and is only used there
for PLY I/O and GUI windows, never for arithmetic on the label-fusion path)."""
