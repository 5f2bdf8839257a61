"""TEST INFRASTRUCTURE ONLY (see oracle/flow_oracle.py header) -- numpy fp64 restatement of the flow simulation that
labels the reference's training data (fluid_generator/generate_xiao_fluid_flow.cpp:147-188, 209-365): D2Q9 BGK
lattice-Boltzmann channel flow, nu = u_max*200/Re (tau = 3 nu + 1/2 = 0.65), parabolic Zou-He velocity inlet on the
left column (:147-165), Zou-He density outlet (rho = 1) on the right column (:170-187), bounce-back solids (objects +
top/bottom walls, :345-352), uniform start at rho = 1, v = (0.08, 0) (:355-363), 10000 steps (:365).

PARITY UNPINNED: the time loop itself lives in MechSys (`mechsys/flbm/Domain.h`, un-vendored, version unknown; SURVEY
3.5).  This restates the textbook scheme with the conventions visible in the reference file: velocity set
C = (0,0),(1,0),(0,1),(-1,0),(0,-1),(1,1),(-1,1),(-1,-1),(1,-1) (the Zou-He lines pair 1<->3, 5<->7, 8<->6), per step
boundary columns -> macroscopic fields -> collide (solids: full bounce-back) -> periodic streaming.
Arrays are [B, ny, nx] (the .h5 files are iy-major: utils/flow_reader.py reshapes to [ny, nx, 3])."""
import numpy as np

CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])
WK = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)
OPP = np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])


def inlet_profile(ny, u_max):
    """:239-246"""
    i = np.arange(ny, dtype=np.float64)
    L = ny - 2.0
    yp = i - 1.5
    return u_max * 4 / (L * L) * (L * yp - yp * yp)


def feq(rho, vx, vy):
    cu = CX[:, None, None, None] * vx[None] + CY[:, None, None, None] * vy[None]
    return WK[:, None, None, None] * rho[None] * (1 + 3 * cu + 4.5 * cu * cu - 1.5 * (vx * vx + vy * vy)[None])


def init(B, ny, nx, rho0=1.0, v0=(0.08, 0.0)):
    rho = np.full((B, ny, nx), rho0)
    return feq(rho, np.full((B, ny, nx), v0[0]), np.full((B, ny, nx), v0[1]))          # [9, B, ny, nx]


def apply_bc(f, vin, rho_out):
    """Zou-He columns, applied to every row as the reference does (:150-157, :174-179).  f: [9,B,ny,nx], in place."""
    v = vin[None, :]
    l = f[:, :, :, 0]
    rho = (l[0] + l[2] + l[4] + 2.0 * (l[3] + l[6] + l[7])) / (1.0 - v)
    l[1] = l[3] + (2.0 / 3.0) * rho * v
    l[5] = l[7] + (1.0 / 6.0) * rho * v - 0.5 * (l[2] - l[4])
    l[8] = l[6] + (1.0 / 6.0) * rho * v + 0.5 * (l[2] - l[4])
    r = f[:, :, :, -1]
    vx = -1.0 + (r[0] + r[2] + r[4] + 2.0 * (r[1] + r[5] + r[8])) / rho_out
    r[3] = r[1] - (2.0 / 3.0) * rho_out * vx
    r[7] = r[5] - (1.0 / 6.0) * rho_out * vx + 0.5 * (r[2] - r[4])
    r[6] = r[8] - (1.0 / 6.0) * rho_out * vx - 0.5 * (r[2] - r[4])


def macroscopic(f, solid):
    rho = f.sum(0)
    safe = np.where(solid, 1.0, rho)
    vx = (CX[:, None, None, None] * f).sum(0) / safe
    vy = (CY[:, None, None, None] * f).sum(0) / safe
    return np.where(solid, 0.0, rho), np.where(solid, 0.0, vx), np.where(solid, 0.0, vy)


def step(f, solid, vin, tau, rho_out=1.0):
    """One time step; returns the new populations."""
    f = f.copy()
    apply_bc(f, vin, rho_out)
    rho, vx, vy = macroscopic(f, solid)
    post = f - (f - feq(np.where(solid, 1.0, rho), vx, vy)) / tau
    post = np.where(solid[None], f[OPP], post)
    out = np.empty_like(f)
    for k in range(9):
        out[k] = np.roll(np.roll(post[k], CY[k], axis=1), CX[k], axis=2)
    return out


def run(solid, steps, u_max=0.1, Re=400.0, rho_out=1.0):
    """solid: bool [B, ny, nx].  Returns (rho, vx, vy) [B, ny, nx] after `steps` steps (boundary columns closed)."""
    B, ny, nx = solid.shape
    tau = 3.0 * (u_max * 200.0 / Re) + 0.5
    vin = inlet_profile(ny, u_max)
    f = init(B, ny, nx)
    for _ in range(steps):
        f = step(f, solid, vin, tau, rho_out)
    apply_bc(f, vin, rho_out)
    return macroscopic(f, solid)
