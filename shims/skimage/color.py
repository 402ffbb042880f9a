def rgb2lab(*a, **k):
    raise NotImplementedError("skimage shim: colour conversion is part of the reference's evaluation tooling, out of scope")
