"""A minimal, scalar, float32 interpreter of the Warp constructs the reference's kernels use,
so that their *unmodified Python source* can be executed thread by thread on the CPU.
TEST INFRASTRUCTURE for generating golden vectors only (tests/golden/make_golden_mpm.py).

Semantics followed (warp-lang 0.6.1 docs): float literals are float32; int() truncates toward
zero; wp.mat33(a, b, c) built from three vec3 takes them as COLUMNS (the reference relies on it,
mpm.py:350 "wp.mat33(col_vec, col_vec, col_vec)"); wp.mat33 of nine scalars is row-major;
mat * vec is the matrix-vector product, mat * mat the matrix product; wp.svd3 returns U, sigma, V
with A = U diag(sigma) V^T (numpy's LAPACK SVD stands in for Warp's McAdams routine)."""
import sys
import types

import numpy as np

f32 = np.float32
_tid = [0]


class vec3:
    __slots__ = ("a",)

    def __init__(self, *args):
        if len(args) == 1 and isinstance(args[0], np.ndarray):
            self.a = args[0].astype(f32)
        elif len(args) == 1:
            self.a = np.full(3, args[0], f32)
        elif len(args) == 0:
            self.a = np.zeros(3, f32)
        else:
            self.a = np.array(args, f32)

    def __getitem__(self, i):
        return self.a[i]

    def __add__(self, o):
        return vec3(self.a + o.a)

    def __sub__(self, o):
        return vec3(self.a - o.a)

    def __mul__(self, s):
        return vec3(self.a * f32(s))

    __rmul__ = __mul__

    def __truediv__(self, s):
        return vec3(self.a / f32(s))

    def __neg__(self):
        return vec3(-self.a)


class mat33:
    __slots__ = ("m",)

    def __init__(self, *args):
        if len(args) == 1 and isinstance(args[0], np.ndarray):
            self.m = args[0].astype(f32)
        elif len(args) == 3:
            self.m = np.stack([a.a for a in args], axis=1).astype(f32)   # columns
        elif len(args) == 9:
            self.m = np.array(args, f32).reshape(3, 3)                    # row-major
        else:
            raise TypeError("mat33 shim: unsupported constructor")

    def __getitem__(self, ij):
        return self.m[ij[0], ij[1]]

    def __add__(self, o):
        return mat33(self.m + o.m)

    def __sub__(self, o):
        return mat33(self.m - o.m)

    def __mul__(self, o):
        if isinstance(o, mat33):
            r = np.zeros((3, 3), f32)
            for i in range(3):
                for j in range(3):
                    acc = f32(0)
                    for k in range(3):
                        acc = f32(acc + self.m[i, k] * o.m[k, j])
                    r[i, j] = acc
            return mat33(r)
        if isinstance(o, vec3):
            r = np.zeros(3, f32)
            for i in range(3):
                acc = f32(0)
                for k in range(3):
                    acc = f32(acc + self.m[i, k] * o.a[k])
                r[i] = acc
            return vec3(r)
        return mat33(self.m * f32(o))

    def __rmul__(self, s):
        return mat33(self.m * f32(s))


class Arr:
    """wp.array of scalars / vec3 / mat33 over a numpy buffer."""

    def __init__(self, buf, kind):
        self.buf, self.kind = buf, kind
        self.shape = buf.shape[:buf.ndim - {"s": 0, "v": 1, "m": 2}[kind]]

    def __getitem__(self, idx):
        x = self.buf[idx]
        return x if self.kind == "s" else (vec3(x.copy()) if self.kind == "v" else mat33(x.copy()))

    def __setitem__(self, idx, val):
        self.buf[idx] = val if self.kind == "s" else (val.a if self.kind == "v" else val.m)


def install():
    wp = types.ModuleType("warp")
    wp.vec3, wp.mat33 = vec3, mat33
    wp.float32, wp.int32 = f32, np.int32
    wp.kernel = lambda fn: fn
    wp.struct = lambda cls: cls
    wp.array = lambda *a, **k: None
    wp.tid = lambda: _tid[0]
    wp.Tape = type("Tape", (), {})
    wp.cw_mul = lambda a, b: vec3(a.a * b.a)
    wp.outer = lambda a, b: mat33(np.outer(a.a, b.a).astype(f32))
    wp.transpose = lambda m: mat33(m.m.T.copy())
    wp.determinant = lambda m: f32(np.linalg.det(m.m.astype(np.float64)))
    wp.clamp = lambda x, lo, hi: min(max(f32(x), f32(lo)), f32(hi))

    def atomic_add(arr, *args):
        idx, val = tuple(args[:-1]), args[-1]
        if arr.kind == "v":
            arr.buf[idx] = (arr.buf[idx] + val.a).astype(f32)
        else:
            arr.buf[idx] = f32(arr.buf[idx] + f32(val))
    wp.atomic_add = atomic_add

    def svd3(A, U, s, V):
        u, sv, vh = np.linalg.svd(A.m.astype(np.float64))
        U.m[:] = u.astype(f32)
        s.a[:] = sv.astype(f32)
        V.m[:] = vh.T.astype(f32)
    wp.svd3 = svd3
    ctx = types.SimpleNamespace(Devicelike=object)
    wp.context = ctx
    wp.codegen = types.SimpleNamespace(StructInstance=object)
    sys.modules["warp"] = wp
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    sys.modules["omegaconf"] = oc
    return wp


def launch(kernel, dim, inputs):
    """wp.launch(kernel, dim=..., inputs=[...]) -- serial over the thread index."""
    if isinstance(dim, int):
        for t in range(dim):
            _tid[0] = t
            kernel(*inputs)
    else:
        for a in range(dim[0]):
            for b in range(dim[1]):
                for c in range(dim[2]):
                    _tid[0] = (a, b, c)
                    kernel(*inputs)
