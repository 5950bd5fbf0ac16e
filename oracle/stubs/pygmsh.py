"""Stand-in for pygmsh (absent from this image); oracle test infrastructure only.
Any use raises: the hot path never touches it."""


def __getattr__(name):
    raise ImportError("pygmsh is not available in this image (oracle stub): " + name)
