"""ctypes binding of libneuma_b200.so (C ABI: include/neuma_b200.h).

There is no fallback: if the shared library is missing or the device is not
sm_100, every entry point raises.  The library is built in-tree by
``__graft_entry__.build()`` / ``python -m neuma_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# NEUMA_B200_LIB: another build of the same library (A/B tooling: tools/gather_variants.sh)
LIB_PATH = os.environ.get("NEUMA_B200_LIB") or os.path.join(_HERE, "libneuma_b200.so")


class MPMConstC(C.Structure):
    """neuma_mpm_const (MPMConstant, modules/nclaw/sim/mpm.py:158-167)."""
    _fields_ = [
        ("num_grids", C.c_int32), ("dt", C.c_float), ("bound", C.c_int32),
        ("gravity", C.c_float * 3), ("dx", C.c_float), ("inv_dx", C.c_float),
        ("eps", C.c_float), ("bc", C.c_int32), ("z_off", C.c_int32), ("nz", C.c_int32),
        ("compact", C.c_int32), ("cap_nodes", C.c_int32),
        ("sel_mode", C.c_int32), ("sel_lo_end", C.c_int32), ("sel_up_start", C.c_int32),
    ]

    def variant(self, **kw) -> "MPMConstC":
        """A copy with some fields replaced (the struct is passed by value per call)."""
        c = MPMConstC.from_buffer_copy(self)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


class HaloSideC(C.Structure):
    """neuma_halo_side: one neighbour of the Z-slab halo exchange."""
    _fields_ = [
        ("planes", C.c_void_p), ("peer_mail", C.c_void_p), ("peer_hdr", C.c_void_p),
        ("my_mail", C.c_void_p), ("my_hdr", C.c_void_p), ("peer_data_flag", C.c_void_p),
        ("my_data_flag", C.c_void_p), ("peer_ack_flag", C.c_void_p), ("my_ack_flag", C.c_void_p),
        ("npl", C.c_int32), ("z0", C.c_int32),
    ]


class RasterSettingsC(C.Structure):
    """neuma_raster_settings (GaussianRasterizationSettings)."""
    _fields_ = [
        ("image_height", C.c_int32), ("image_width", C.c_int32),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("bg", C.c_float * 3),
        ("scale_modifier", C.c_float), ("viewmatrix", C.c_float * 16),
        ("projmatrix", C.c_float * 16), ("sh_degree", C.c_int32), ("campos", C.c_float * 3),
        ("prefiltered", C.c_int32), ("debug", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_CP = C.POINTER(MPMConstC)
_RP = C.POINTER(RasterSettingsC)

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "neuma_ctx_create": [_I, C.POINTER(_P)],
    "neuma_ctx_destroy": [_P],
    "neuma_last_error": [_P],
    "neuma_version": [],
    "neuma_launch_count": [_P],
    "neuma_mpm_grid_clear": [_P, _CP, _P],
    "neuma_mpm_p2g": [_P, _CP, _I] + [_P] * 8,
    "neuma_mpm_grid_op": [_P, _CP, _P],
    "neuma_mpm_g2p": [_P, _CP, _I] + [_P] * 8 + [_I, _P],
    "neuma_mpm_grid_export": [_P, _CP, _P, _P, _P, _P],
    "neuma_mpm_halo_pack": [_P, _CP, _I, _I, _P, _P],
    "neuma_mpm_halo_unpack_add": [_P, _CP, _I, _I, _P, _P],
    "neuma_mpm_substep_fwd": [_P, _CP, _I] + [_P] * 14,
    "neuma_mpm_substep_bwd": [_P, _CP, _I] + [_P] * 20,
    "neuma_mpm_box_from_x": [_P, _CP, _I, _P, _P, _P, _P],
    "neuma_mpm_box_clear": [_P, _CP, _P, _P, _P, _P, _P],
    "neuma_mpm_p2g_to": [_P, _CP, _I] + [_P] * 10,
    "neuma_mpm_grid_op_box": [_P, _CP, _P, _P, _P, _P],
    "neuma_mpm_bwd_node": [_P, _CP, _P, _P, _P, _P],
    "neuma_mpm_grid_op_next": [_P, _CP, _P, _P, _P, _P, _P, _P],
    "neuma_mpm_g2p_from": [_P, _CP, _I] + [_P] * 12,
    "neuma_mpm_rollout_fwd": [_P, _CP, _I] + [_P] * 18,
    "neuma_mpm_rollout_bwd": [_P, _CP, _I] + [_P] * 24,
    "neuma_mpm_bwd_scatter": [_P, _CP, _I] + [_P] * 12,
    "neuma_mpm_bwd_node_gather": [_P, _CP, _I] + [_P] * 24,
    "neuma_mpm_planes_add": [_P, _P, _P, C.c_size_t, _P],
    "neuma_halo_push": [_P, _P, _P, C.c_size_t, _P, _P, _P, _P, _P],
    "neuma_halo_add_ack": [_P, _P, _P, C.c_size_t, _P, _P, _P, _P, _P],
    "neuma_halo_exchange": [_P, _CP, C.POINTER(HaloSideC), C.POINTER(HaloSideC), _P, _P, _P],
    "neuma_mpm_cell_order": [_P, _CP, _I, _P, _P, _P],
    "neuma_gather_rows": [_P, _I, _I, _P, _P, _P, _P],
    "neuma_invert_perm": [_P, _I, _P, _P, _P],
    "neuma_svd3_fwd": [_P, _I, _P, _P, _P, _P, _P],
    "neuma_svd3_bwd": [_P, _I] + [_P] * 8,
    "neuma_constitutive_fwd": [_P, _I, _I, _I, _P, _P, _P, _P, _F, _P, _P],
    "neuma_constitutive_bwd": [_P, _I, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P],
    "neuma_constitutive_staged_bytes": [],
    "neuma_constitutive_stage": [_P, _I, _P, _P, _P, _P, _P],
    "neuma_constitutive_fwd_staged": [_P, _I, _I, _I, _P, _P, _F, _P, _P],
    "neuma_constitutive_bwd_staged": [_P, _I, _I, _I, _P, _P, _F, _P, _P, _P, _P, _P, _P],
    "neuma_constitutive_partials_floats": [_P],
    "neuma_constitutive_bwd_partials": [_P, _I, _I, _I, _P, _P, _F, _P, _P, _I, _P, _P],
    "neuma_constitutive_fwd_p2g": [_P, _I, _I, _I, _P, _P, _P, _CP] + [_P] * 9,
    "neuma_constitutive_bwd_scatter": [_P, _I, _I, _P, _P, _F, _P, _P, _P, _CP] + [_P] * 11,
    "neuma_constitutive_fwd_g2p": [_P, _I, _I, _P, _F, _P, _CP] + [_P] * 12,
    "neuma_constitutive_bwd_gather": [_P, _I, _I, _I, _P, _P, _CP] + [_P] * 24,
    "neuma_constitutive_reduce_partials": [_P, _I, _P, _P, _P, _P, _P],
    "neuma_constitutive_svd_cache": [_P, _P, _P],
    "neuma_bind_spmm": [_P, _I, _I, _P, _P, _P, _I, _P, _P, _I, _P],
    "neuma_particles_to_gaussians": [_P, _I] + [_P] * 15,
    "neuma_gaussians_to_particles": [_P, _I] + [_P] * 8,
    "neuma_bind_build": [_P, _I, _I, _P, _P, _P, _F, _I, _P, _P, _P, _P],
    "neuma_deform_cov": [_P, _I, _P, _P, _P, _P],
    "neuma_raster_fwd": [_P, _RP, _I, _I] + [_P] * 8,
    "neuma_raster_bwd": [_P, _RP, _I, _I] + [_P] * 8 + [_I] + [_P] * 6,
    "neuma_raster_state_bytes": [_P, C.POINTER(C.c_size_t)],
    "neuma_raster_state_save": [_P, _P, _P],
    "neuma_raster_num_rendered": [_P, _P, C.POINTER(_I)],
    "neuma_raster_export_binning": [_P, _P, _P, _P, _P],
}
_RESTYPES = {"neuma_last_error": C.c_char_p, "neuma_launch_count": C.c_longlong}

_lib: Optional[C.CDLL] = None


class NeumaError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libneuma_b200.so; loud failure if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NeumaError(
            f"{LIB_PATH} not found: build it with `python -m neuma_b200.build` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: loud by design
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


class Context:
    """One neuma_ctx per CUDA device (owns the grid / sort workspaces)."""

    _by_device: dict = {}

    def __init__(self, device: int):
        self.lib = load()
        self.device = int(device)
        h = _P()
        rc = self.lib.neuma_ctx_create(self.device, C.byref(h))
        if rc != 0:
            why = {3: "no such CUDA device", 4: "device is not sm_100 (B200): no fallback path"}.get(rc, "")
            raise NeumaError(f"neuma_ctx_create(device={device}) failed with code {rc} {why}")
        self.handle = h

    @classmethod
    def get(cls, device: int) -> "Context":
        device = int(device)
        ctx = cls._by_device.get(device)
        if ctx is None:
            ctx = cls._by_device[device] = cls(device)
        return ctx

    def call(self, name: str, *args) -> None:
        import torch
        if torch.cuda.current_device() != self.device:          # allocations / launches of the library
            with torch.cuda.device(self.device):                # belong to the ctx's device
                rc = getattr(self.lib, name)(self.handle, *args)
        else:
            rc = getattr(self.lib, name)(self.handle, *args)
        if rc != 0:
            msg = self.lib.neuma_last_error(self.handle)
            raise NeumaError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def launches(self) -> int:
        return int(self.lib.neuma_launch_count(self.handle))

    def close(self) -> None:
        if self.handle:
            self.lib.neuma_ctx_destroy(self.handle)
            self.handle = None


def ptr(t) -> C.c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return _P(0)
    return _P(t.data_ptr())
