"""TEST INFRASTRUCTURE ONLY -- stand-in for scikit-spatial (absent from this image), needed to import the reference's
`Fusion3DSeg/merge_intersecting_bb.py:11` unmodified.  Only `Line.project_point` is used (`:20-36`, `cal_min_max`)."""
