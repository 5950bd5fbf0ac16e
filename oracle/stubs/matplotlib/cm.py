from matplotlib import _Inert


def __getattr__(name):
    return _Inert()
