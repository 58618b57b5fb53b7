"""ctypes view of oracle/conv_ref.c (plain-C conv primitives).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libconv_ref.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


def _lib():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "conv_ref.c")):
        build()
    return ctypes.CDLL(_SO)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def conv1d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
    x, w = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(w, np.float32)
    b, cin, t = x.shape
    cout, _, k = w.shape
    tout = (t + 2 * padding - dilation * (k - 1) - 1) // stride + 1
    y = np.zeros((b, cout, tout), np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    rc = _lib().ref_conv1d(_p(x), _p(w), _p(bias), _p(y), b, cin, t, cout, k, stride, padding, dilation, groups)
    assert rc == tout
    return y


def conv_transpose1d(x, w, bias=None, stride=1, padding=0):
    x, w = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(w, np.float32)
    b, cin, t = x.shape
    _, cout, k = w.shape
    tout = (t - 1) * stride - 2 * padding + k
    y = np.zeros((b, cout, tout), np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    rc = _lib().ref_conv_transpose1d(_p(x), _p(w), _p(bias), _p(y), b, cin, t, cout, k, stride, padding)
    assert rc == tout
    return y


def avg_pool1d(x, k, stride, padding):
    x = np.ascontiguousarray(x, np.float32)
    b, c, t = x.shape
    tout = (t + 2 * padding - k) // stride + 1
    y = np.zeros((b, c, tout), np.float32)
    rc = _lib().ref_avg_pool1d(_p(x), _p(y), b * c, t, k, stride, padding)
    assert rc == tout
    return y
