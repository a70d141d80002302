"""Stand-in for scikit-image (absent from this image): the reference's `RTAB_utils/ios_rtab.py` imports
`skimage.transform.resize` at module level but only calls it for RGB colours, which the label-fusion path never reads."""
