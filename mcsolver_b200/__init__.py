"""mcsolver_b200 - B200-native Monte Carlo engine behind the native-library boundary of golddoushi/mcsolver.

    import mcsolver_b200
    mcsolver_b200.loadMC("parameterfile")        # the reference's headless entry, on the GPU

No PyTorch in the product path; the CUDA library is built in-tree by `python -m mcsolver_b200.build`.
"""
__version__ = "0.1.0"


def loadMC(rpath, **kw):
    from .driver import loadMC as _l
    return _l(rpath, **kw)
