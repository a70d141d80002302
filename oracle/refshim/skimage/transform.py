"""See the package docstring: importable, not callable."""


def resize(*args, **kwargs):
    raise NotImplementedError("skimage.transform.resize stand-in: RGB colour resampling is outside the label-fusion path")
