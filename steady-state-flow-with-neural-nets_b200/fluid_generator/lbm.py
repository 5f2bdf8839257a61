"""GPU flow simulation for labelled training data: the D2Q9 BGK lattice-Boltzmann channel flow of
/root/reference/fluid_generator/generate_xiao_fluid_flow.cpp, which needs the un-vendored MechSys library, as kernels
of libflownet.so (csrc/lbm.cu).  B independent simulations advance together; the lattices stay on the device."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

import flownet_native as native

# reference constants (generate_xiao_fluid_flow.cpp:222-231, 355-365)
U_MAX = 0.1
RE = 400.0
NU = U_MAX * 200.0 / RE
TAU = 3.0 * NU + 0.5
RHO0, V0 = 1.0, (0.08, 0.0)
STEPS = 10000


def domain_from_boundary(boundary, pad=128):
    """Embed dataset boundaries uint8 [B,H,W(,1)] in the simulation domain [B, H, W + 128]: the reference simulates
    nx = 2*ny + 128 columns (:224) and keeps the first 2*ny (utils/flow_reader.py:8); top and bottom rows are walls
    (:345-352)."""
    b = np.asarray(boundary)
    if b.ndim == 4:
        b = b[..., 0]
    B, H, W = b.shape
    solid = np.zeros((B, H, W + pad), np.uint8)
    solid[:, :, :W] = b
    solid[:, 0, :] = 1
    solid[:, H - 1, :] = 1
    return solid


class ChannelFlow:
    def __init__(self, solid, dtype=torch.float32, tau=TAU, u_max=U_MAX, rho_out=1.0, device="cuda"):
        solid = torch.as_tensor(np.ascontiguousarray(solid, dtype=np.uint8))
        self.solid = solid.to(device).contiguous()
        self.B, self.ny, self.nx = self.solid.shape
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("populations are fp32 or fp64")
        self.f_a = torch.empty((self.B, 9, self.ny, self.nx), dtype=dtype, device=device)
        self.f_b = torch.empty_like(self.f_a)
        d = native.LbmDesc()
        d.B, d.nx, d.ny = self.B, self.nx, self.ny
        d.dtype = 0 if dtype == torch.float32 else 1
        d.tau, d.u_max, d.rho_out = float(tau), float(u_max), float(rho_out)
        d.solid, d.f_a, d.f_b = self.solid.data_ptr(), self.f_a.data_ptr(), self.f_b.data_ptr()
        self.desc = d
        self.steps_done = 0
        native._check(native.lib.fn_lbm_init(C.byref(d), RHO0, V0[0], V0[1], native.stream_ptr()), "fn_lbm_init")

    def step(self, nsteps):
        if nsteps % 2:
            raise ValueError("advance by an even number of steps (two-lattice scheme)")
        native._check(native.lib.fn_lbm_steps(C.byref(self.desc), int(nsteps), native.stream_ptr()), "fn_lbm_steps")
        self.steps_done += nsteps

    def fields(self, want_rho=False):
        """velocity fp32 [B, ny, nx, 2] (and density [B, ny, nx]) of the current state."""
        vel = torch.empty((self.B, self.ny, self.nx, 2), dtype=torch.float32, device=self.solid.device)
        rho = torch.empty((self.B, self.ny, self.nx), dtype=torch.float32, device=self.solid.device) if want_rho else None
        native._check(native.lib.fn_lbm_fields(C.byref(self.desc), vel.data_ptr(), rho.data_ptr() if want_rho else None,
                                               native.stream_ptr()), "fn_lbm_fields")
        return (vel, rho) if want_rho else vel


def simulate(boundary, steps=STEPS, dtype=torch.float32):
    """boundary uint8 [B,H,W(,1)] -> (boundary uint8 [B,H,W,1], sflow float32 [B,H,W,2]) as the dataset stores them:
    the steady flow around the obstacles, cropped to the first W columns (utils/flow_reader.py:5-17)."""
    solid = domain_from_boundary(boundary)
    sim = ChannelFlow(solid, dtype=dtype)
    sim.step(steps)
    W = solid.shape[2] - 128
    vel = sim.fields()[:, :, :W, :].contiguous()
    return solid[:, :, :W, None].copy(), vel
