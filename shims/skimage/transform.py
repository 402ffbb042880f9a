def resize(*a, **k):
    raise NotImplementedError("skimage shim: image resizing is part of the reference's evaluation tooling, out of scope")
