"""Camera, TAA jitter and sun direction as the reference's host code computes them (inputs of the trace passes).

  FpsCamera          Core/FpsCamera.cpp:23-24,150-163 (glm::lookAt / glm::perspective, right-handed, -1..1 depth)
  halton_table       Core/TAAJitter.cpp:6-47
  sun_moon_direction Core/Pipeline.cpp:1652-1671
All arithmetic is float32 like glm's; matrices are handed to the C ABI column-major (glm::value_ptr order).
"""
import math

import numpy as np

from .abi import VxCamera

f32 = np.float32


def perspective(fovy_deg, aspect, z_near, z_far):
    t = math.tan(math.radians(fovy_deg) / 2.0)
    m = np.zeros((4, 4), dtype=np.float64)  # m[row, col]
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(z_far + z_near) / (z_far - z_near)
    m[3, 2] = -1.0
    m[2, 3] = -(2.0 * z_far * z_near) / (z_far - z_near)
    return m


def look_at(eye, center, up):
    eye, center, up = (np.asarray(a, dtype=np.float64) for a in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float64)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s.dot(eye), -u.dot(eye), f.dot(eye)
    return m


class FpsCamera:
    """Default player camera: fov 60, near 0.1, far 1000 (Core/Player.cpp:10), start (192, 75, 192) (Pipeline.cpp:1500)."""

    def __init__(self, position=(192.0, 75.0, 192.0), yaw_deg=90.0, pitch_deg=0.0, fov_deg=60.0, aspect=16.0 / 9.0, z_near=0.1, z_far=1000.0):
        self.position = np.asarray(position, dtype=np.float64)
        self.yaw, self.pitch = yaw_deg, pitch_deg
        self.fov, self.aspect, self.z_near, self.z_far = fov_deg, aspect, z_near, z_far

    @property
    def front(self):
        # FPSCamera::UpdateOnMouseMovement, Core/FpsCamera.cpp:66-70 (yaw 90 deg looks down +Z)
        ry, rp = self.yaw * math.pi / 180.0, self.pitch * math.pi / 180.0
        # rounded to float like glm::vec3 front (and like the C++ host mirror)
        return np.array([f32(math.cos(rp) * math.cos(ry)), f32(math.sin(rp)), f32(math.cos(rp) * math.sin(ry))], dtype=np.float64)

    def view(self):
        return look_at(self.position, self.position + self.front, (0.0, 1.0, 0.0))

    def projection(self):
        return perspective(self.fov, self.aspect, self.z_near, self.z_far)

    def inverse_matrices(self):
        """(inv_view, inv_proj) as float32 [4,4] (math convention m[row, col]), computed analytically in double — the same
        formulas, in the same order, as the C++ host mirror (voxelpathtracer_b200/host/VoxelRT.h: FPSCamera::GetVxCamera),
        so both hosts hand bit-identical matrices to the ABI."""
        f = self.front
        f = f / math.sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2])
        s = np.array([f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0])  # cross(f, (0,1,0))
        s = s / math.sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2])
        u = np.array([s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]])  # cross(s, f)
        inv_view = np.zeros((4, 4), dtype=np.float32)
        inv_view[:3, 0], inv_view[:3, 1], inv_view[:3, 2] = s.astype(f32), u.astype(f32), (-f).astype(f32)
        inv_view[:3, 3] = self.position.astype(f32)
        inv_view[3, 3] = 1.0
        t = math.tan(self.fov * math.pi / 180.0 / 2.0)
        A = -(self.z_far + self.z_near) / (self.z_far - self.z_near)
        B = -(2.0 * self.z_far * self.z_near) / (self.z_far - self.z_near)
        inv_proj = np.zeros((4, 4), dtype=np.float32)
        inv_proj[0, 0] = f32(self.aspect * t)
        inv_proj[1, 1] = f32(t)
        inv_proj[3, 2] = f32(1.0 / B)
        inv_proj[2, 3] = -1.0
        inv_proj[3, 3] = f32(A / B)
        return inv_view, inv_proj

    def view_projection_f32(self):
        """(view[16], projection[16]) float32, column-major: u_View / u_Projection of this frame = u_PrevView / u_PrevProjection of the next
        frame's temporal filters.  Same formulas, in the same order, as the C++ host mirror (FPSCamera::GetViewProjection), so both hosts
        hand bit-identical matrices to the ABI (view() / projection() above are the float64 textbook forms)."""
        f = self.front
        f = f / math.sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2])
        s = np.array([f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0])
        s = s / math.sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2])
        u = np.array([s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]])
        e = self.position.astype(f32).astype(np.float64)
        view, proj = np.zeros(16, dtype=np.float32), np.zeros(16, dtype=np.float32)
        for c in range(3):
            view[4 * c + 0], view[4 * c + 1], view[4 * c + 2] = f32(s[c]), f32(u[c]), f32(-f[c])
        view[12] = f32(-((s[0] * e[0] + s[1] * e[1]) + s[2] * e[2]))
        view[13] = f32(-((u[0] * e[0] + u[1] * e[1]) + u[2] * e[2]))
        view[14] = f32((f[0] * e[0] + f[1] * e[1]) + f[2] * e[2])
        view[15] = 1.0
        fov, aspect, zn, zf = float(self.fov), float(self.aspect), float(self.z_near), float(self.z_far)
        t = math.tan(fov * math.pi / 180.0 / 2.0)
        proj[0], proj[5] = f32(1.0 / (aspect * t)), f32(1.0 / t)
        proj[10], proj[11], proj[14] = f32(-(zf + zn) / (zf - zn)), -1.0, f32(-(2.0 * zf * zn) / (zf - zn))
        return view, proj

    def vx_camera(self, width, height, row_begin=0, row_end=None, interleave_n=0, interleave_rank=0, band_rows=0):
        """inv_view / inv_projection as Pipeline.cpp:1823-1824 hands them to the shaders.  interleave_*: see VxCamera."""
        cam = VxCamera()
        inv_view, inv_proj = self.inverse_matrices()
        cam.inv_view[:] = inv_view.T.reshape(16).tolist()  # column-major
        cam.inv_proj[:] = inv_proj.T.reshape(16).tolist()
        cam.width, cam.height = int(width), int(height)
        cam.row_begin = int(row_begin)
        rows = height // interleave_n if interleave_n > 1 else height
        cam.row_end = int(rows if row_end is None else row_end)
        cam.interleave_n, cam.interleave_rank, cam.band_rows = int(interleave_n), int(interleave_rank), int(band_rows)
        return cam


def _halton(prime, index):
    r, f, i = f32(0.0), f32(1.0), index
    while i > 0:
        f = f32(f / f32(prime))
        r = f32(r + f32(f * f32(i % prime)))
        i = int(math.floor(i / float(prime)))
    return float(r)


HALTON_TABLE = [(_halton(2, i + 1), _halton(3, i + 1)) for i in range(64)]  # GenerateJitterStuff


def taa_jitter(frame):
    """GetTAAJitter: entry frame % 64 (Core/TAAJitter.cpp:37-41)."""
    return HALTON_TABLE[frame % 64]


def taa_jitter_secondary(frame):
    """GetTAAJitterSecondary: entry frame % 32 (Core/TAAJitter.cpp:43-47)."""
    return HALTON_TABLE[frame % 32]


def sun_moon_direction(sun_tick=50.0):
    """Core/Pipeline.cpp:1652-1671: rotate (1,1,1) by 2*SunTick degrees about +Z, normalise; moon = (-x,-y,z)."""
    a = math.radians(sun_tick * 2.0)
    c, s = math.cos(a), math.sin(a)
    sun = np.array([c - s, s + c, 1.0])
    moon = np.array([-sun[0], -sun[1], sun[2]])
    sun = sun / np.linalg.norm(sun)
    moon = moon / np.linalg.norm(moon)
    stronger = sun if -sun[1] < 0.01 else moon
    sun_visibility = min(max(float(sun[1]) + 0.05, 0.0), 0.1) * 12.0  # Pipeline.cpp:1828
    return sun.astype(f32), moon.astype(f32), stronger.astype(f32), float(f32(sun_visibility))
