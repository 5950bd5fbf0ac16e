"""Inert stand-in for matplotlib (absent from this image).

TEST INFRASTRUCTURE ONLY: lets `import FEMOL` (the read-only reference under
/root/reference) succeed so its numpy/scipy arithmetic can be executed as the
parity oracle.  Every attribute resolves to an object that swallows calls.
"""


class _Inert(object):
    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Inert()
