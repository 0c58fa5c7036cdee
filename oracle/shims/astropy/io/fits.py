from oracle.fitsio import getdata as _getdata


def getdata(filename, header=False, **kw):
    return _getdata(filename, with_header=header)


def writeto(*a, **k):
    raise NotImplementedError("oracle shim: FITS writing is out of scope")
