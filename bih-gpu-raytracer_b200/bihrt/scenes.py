"""Deterministic synthetic scenes and cameras for the BASELINE.json configs.

The reference ships no meshes (`*.obj` is git-ignored, /root/reference/.gitignore:77), so every
workload is procedural.  All generators return float32 arrays of shape (N, 9): v0.xyz v1.xyz v2.xyz
per triangle, the layout of the reference's `Triangle` (R/src/Tree.cuh:37-46), in the order the
reference's LoadModels would see them (model -> mesh -> face, R/src/App.cpp:110-121).
Seeds: integer hash only, seed 1984 (echoing R/src/CUDAKernels.cu:458).
"""
import numpy as np

SEED = 1984


def _mix32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def _hash01(i, j=0, seed=SEED):
    h = _mix32(np.asarray(i, np.uint64) * 0x9E3779B1 + np.asarray(j, np.uint64) * 0x85EBCA77 + seed)
    return (h >> 8).astype(np.float64) / 16777216.0


# ---------------------------------------------------------------------------------------------
# known-answer mesh: SURVEY.md Appendix A (reconstructed from R/BIH1.txt)
# ---------------------------------------------------------------------------------------------
def dodecahedron(rounded=True):
    if rounded:
        a, b, c = 0.57735, 0.356822, 0.934172
    else:
        phi = (1 + 5 ** 0.5) / 2
        a = 1 / 3 ** 0.5
        b, c = a / phi, a * phi
    v = np.array([
        [-a, -a, -a], [-a, -a, a], [-a, a, -a], [-a, a, a], [a, -a, -a], [a, -a, a], [a, a, -a], [a, a, a],
        [0, -c, -b], [-c, -b, 0], [-b, 0, -c], [0, -c, b], [-c, b, 0], [-b, 0, c], [0, c, -b], [c, -b, 0],
        [b, 0, -c], [0, c, b], [c, b, 0], [b, 0, c]], dtype=np.float32)
    f = np.array([
        (9, 0, 8), (9, 8, 11), (9, 11, 1), (12, 2, 10), (12, 10, 0), (12, 0, 9), (16, 4, 8), (16, 8, 0),
        (16, 0, 10), (12, 9, 1), (12, 1, 13), (12, 13, 3), (11, 5, 19), (11, 19, 13), (11, 13, 1),
        (12, 3, 17), (12, 17, 14), (12, 14, 2), (14, 6, 16), (14, 16, 10), (14, 10, 2), (17, 3, 13),
        (17, 13, 19), (17, 19, 7), (15, 5, 11), (15, 11, 8), (15, 8, 4), (18, 15, 4), (18, 4, 16),
        (18, 16, 6), (18, 7, 19), (18, 19, 5), (18, 5, 15), (18, 6, 14), (18, 14, 17), (18, 17, 7)])
    return np.ascontiguousarray(v[f].reshape(-1, 9))


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def quad_grid(p0, du, dv, nu, nv):
    """Tessellated parallelogram p0 + s*du + t*dv, 2*nu*nv triangles, normal = du x dv."""
    p0, du, dv = (np.asarray(x, np.float64) for x in (p0, du, dv))
    s = np.arange(nu + 1) / nu
    t = np.arange(nv + 1) / nv
    P = p0[None, None, :] + s[:, None, None] * du[None, None, :] + t[None, :, None] * dv[None, None, :]
    v00, v10, v01, v11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    t1 = np.stack([v00, v10, v11], axis=2)
    t2 = np.stack([v00, v11, v01], axis=2)
    tris = np.stack([t1, t2], axis=2)                 # (nu, nv, 2, 3, 3)
    return tris.reshape(-1, 9).astype(np.float32)


def box(lo, hi, inward=False, open_bottom=False, n=1):
    """Axis-aligned box, 12 (or 10 when open_bottom) * n*n triangles; outward normals unless inward."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    faces = [
        (lo, ez, ey),                  # -x : normal ez x ey = -x
        (lo + ex, ey, ez),             # +x
        (lo + ey, ez, ex),             # +y : ez x ex = +y
        (lo, ey, ex),                  # -z : ey x ex = -z
        (lo + ez, ex, ey),             # +z
    ]
    if not open_bottom:
        faces.append((lo, ex, ez))     # -y : ex x ez = -y
    out = []
    for p0, du, dv in faces:
        out.append(quad_grid(p0, dv, du, n, n) if inward else quad_grid(p0, du, dv, n, n))
    return np.concatenate(out)


def cylinder(base, radius, height, nseg, nh, outward=True, arc=(0.0, 2 * np.pi), axis="y"):
    """Open cylinder (or arc of one) around the y (or x / z) axis."""
    a = arc[0] + (arc[1] - arc[0]) * np.arange(nseg + 1) / nseg
    hh = height * np.arange(nh + 1) / nh
    ca, sa = np.cos(a), np.sin(a)
    P = np.empty((nseg + 1, nh + 1, 3))
    if axis == "y":
        P[..., 0] = base[0] + radius * ca[:, None]
        P[..., 1] = base[1] + hh[None, :]
        P[..., 2] = base[2] + radius * sa[:, None]
    elif axis == "z":
        P[..., 0] = base[0] + radius * ca[:, None]
        P[..., 1] = base[1] + radius * sa[:, None]
        P[..., 2] = base[2] + hh[None, :]
    else:
        P[..., 0] = base[0] + hh[None, :]
        P[..., 1] = base[1] + radius * ca[:, None]
        P[..., 2] = base[2] + radius * sa[:, None]
    v00, v10, v01, v11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    t1 = np.stack([v00, v01, v11], axis=2)
    t2 = np.stack([v00, v11, v10], axis=2)
    tris = np.stack([t1, t2], axis=2).reshape(-1, 3, 3)
    # orientation check on the first triangle, flip all if needed
    n0 = np.cross(tris[0, 1] - tris[0, 0], tris[0, 2] - tris[0, 0])
    c0 = tris[0].mean(0) - np.asarray(base, np.float64)
    if axis == "y":
        c0[1] = 0
    elif axis == "z":
        c0[2] = 0
    else:
        c0[0] = 0
    if (np.dot(n0, c0) > 0) != outward:
        tris = tris[:, [0, 2, 1], :]
    return tris.reshape(-1, 9).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# Config 1: Cornell box, 32 triangles (5 inward walls x2, light quad x2, two open-bottom boxes x10)
# ---------------------------------------------------------------------------------------------
def cornell_box():
    room = box([-1, -1, -1], [1, 1, 1], inward=True)                       # 12 tris: 6 faces
    # drop the front wall (-z, faces the camera): keep 5 walls.  faces are emitted in the order
    # -x, +x, +y, -z, +z, -y -> remove face 3
    room = np.concatenate([room[:6], room[8:]])
    light = quad_grid([-0.25, 0.999, -0.25], [0.5, 0, 0], [0, 0, 0.5], 1, 1)  # normal = x cross z = -y (down)
    tall = box([-0.65, -1.0, 0.1], [-0.15, 0.3, 0.6], open_bottom=True)
    short = box([0.15, -1.0, -0.5], [0.65, -0.4, 0.0], open_bottom=True)
    tris = np.concatenate([room, light, tall, short])
    assert tris.shape[0] == 32
    return np.ascontiguousarray(tris)


# ---------------------------------------------------------------------------------------------
# Configs 2/3(fallback)/4/5: displaced UV sphere, 2*nseg^2 triangles, outward CCW
#   radius = 1 + 0.05*(sin(7*theta + phase)*cos(5*phi) + 0.3*hash(i,j))
# nseg=187 -> 69 938, 361 -> 260 642, 708 -> 1 002 528, 2236 -> 9 999 392
# ---------------------------------------------------------------------------------------------
def displaced_sphere(nseg, phase=0.0, seed=SEED):
    i = np.arange(nseg + 1)
    j = np.arange(nseg)
    theta = np.pi * i / nseg
    phi = 2 * np.pi * j / nseg
    h = _hash01(i[:, None], j[None, :], seed)
    r = 1.0 + 0.05 * (np.sin(7 * theta + phase)[:, None] * np.cos(5 * phi)[None, :] + 0.3 * h)
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    st[0, 0] = 0.0
    st[-1, 0] = 0.0
    # poles collapse to a single point so the mesh is closed (pole triangles are degenerate)
    r[0, :] = r[0, 0]
    r[-1, :] = r[-1, 0]
    P = np.stack([r * st * np.cos(phi)[None, :], r * ct * np.ones_like(phi)[None, :],
                  r * st * np.sin(phi)[None, :]], axis=-1).astype(np.float32)   # (nseg+1, nseg, 3)
    Pn = np.concatenate([P, P[:, :1]], axis=1)                                   # seam shares vertices
    v00, v01, v10, v11 = Pn[:-1, :-1], Pn[:-1, 1:], Pn[1:, :-1], Pn[1:, 1:]
    t1 = np.stack([v00, v01, v10], axis=2)
    t2 = np.stack([v01, v11, v10], axis=2)
    tris = np.stack([t1, t2], axis=2)                                           # (nseg, nseg, 2, 3, 3)
    return np.ascontiguousarray(tris.reshape(-1, 9))


SPHERE_NSEG = {"70k": 187, "260k": 361, "1m": 708, "10m": 2236}


# ---------------------------------------------------------------------------------------------
# Config 3: procedural atrium ("Sponza-scale", ~260 k triangles)
# ---------------------------------------------------------------------------------------------
def atrium(detail=0.89):
    """Hall 4 x 2 x 2 (x,y,z) around the origin: tessellated floor / ceiling / walls, two rows of
    columns, barrel-vault arches between them.  detail=0.89 (the default) gives ~260 k triangles."""
    s = max(detail, 1e-3) ** 0.5
    g = lambda n: max(1, int(round(n * s)))
    parts = []
    L, H, W = 2.0, 1.0, 1.0                                   # half extents
    parts.append(quad_grid([-L, -H, -W], [0, 0, 2 * W], [2 * L, 0, 0], g(120), g(240)))   # floor, normal +y
    parts.append(quad_grid([-L, H, -W], [2 * L, 0, 0], [0, 0, 2 * W], g(160), g(80)))     # ceiling, normal -y
    parts.append(quad_grid([-L, -H, W], [0, 2 * H, 0], [2 * L, 0, 0], g(80), g(160)))     # back wall z=+W, normal -z
    parts.append(quad_grid([-L, -H, -W], [2 * L, 0, 0], [0, 2 * H, 0], g(160), g(80)))    # front wall z=-W, normal +z
    parts.append(quad_grid([-L, -H, -W], [0, 2 * H, 0], [0, 0, 2 * W], g(80), g(80)))     # left wall x=-L, normal +x
    parts.append(quad_grid([L, -H, -W], [0, 0, 2 * W], [0, 2 * H, 0], g(80), g(80)))      # right wall x=+L, normal -x
    ncol = 8
    xs = -L + (np.arange(ncol) + 0.5) * (2 * L / ncol)
    for zrow in (-0.45, 0.45):
        for x in xs:
            parts.append(cylinder([x, -H, zrow], 0.07, 1.3, g(48), g(72)))
            parts.append(box([x - 0.1, -H, zrow - 0.1], [x + 0.1, -H + 0.06, zrow + 0.1], n=g(6)))
            parts.append(box([x - 0.09, 0.3, zrow - 0.09], [x + 0.09, 0.36, zrow + 0.09], n=g(6)))
    # barrel-vault arches spanning neighbouring columns along x (seen from below: inward-facing)
    for zrow in (-0.45, 0.45):
        for k in range(ncol - 1):
            xc = 0.5 * (xs[k] + xs[k + 1])
            rad = 0.5 * (xs[k + 1] - xs[k]) - 0.07
            parts.append(cylinder([xc, 0.36, zrow - 0.08], rad, 0.16, g(40), g(8), outward=False,
                                  arc=(0.0, np.pi), axis="z"))
    return np.ascontiguousarray(np.concatenate(parts))


def random_soup(n, size=0.02, seed=SEED):
    idx = np.arange(n)
    c = np.stack([_hash01(idx, 1, seed), _hash01(idx, 2, seed), _hash01(idx, 3, seed)], -1) * 2 - 1
    o = np.stack([_hash01(idx, 4 + k, seed) for k in range(9)], -1).reshape(n, 3, 3) * 2 - 1
    return np.ascontiguousarray((c[:, None, :] + size * o).reshape(n, 9).astype(np.float32))


# ---------------------------------------------------------------------------------------------
# cameras: 12 floats = origin, lowerLeftCorner, horizontal, vertical (R/src/Camera.h:14-17)
# ---------------------------------------------------------------------------------------------
def reference_camera(origin=(2.0, 0.0, -2.0), aspect=640.0 / 480.0):
    """Camera::Camera, R/src/Camera.cu:5-9; default origin from R/src/Renderer.cpp:99."""
    o = np.asarray(origin, np.float32)
    llc = np.array([np.float32(float(o[0]) - 2.0), np.float32(float(o[1]) - 1.0),
                    np.float32(float(o[2]) + 1.0)], np.float32)
    hor = np.array([np.float32(aspect * 2.0), 0, 0], np.float32)
    ver = np.array([0, 2.0, 0], np.float32)
    return np.concatenate([o, llc, hor, ver]).astype(np.float32)


def pinhole_camera(origin=(0.1, 0.2, -3.0), half=0.45, aspect=1920.0 / 1080.0):
    """SURVEY.md 8(d) Config 2: looking +z, unnormalised dirs ((2u-1)*half*aspect, (2v-1)*half, 1)."""
    o = np.asarray(origin, np.float64)
    llc = o + np.array([-half * aspect, -half, 1.0])
    hor = np.array([2 * half * aspect, 0, 0])
    ver = np.array([0, 2 * half, 0])
    return np.concatenate([o, llc, hor, ver]).astype(np.float32)


def cornell_camera():
    # slightly off-axis so pixel centres do not fall exactly on the walls' quad diagonals (a systematic
    # edge-tie case for a perfectly symmetric view: both triangles of a quad report the hit)
    return pinhole_camera(origin=(0.0137, 0.0211, -3.4), half=0.42, aspect=1.0)


def atrium_camera(aspect=1920.0 / 1080.0):
    # inside the hall near the left end, looking along +x:  generic frame from look-at
    return look_at_camera((-1.8, -0.35, 0.0), (1.0, -0.1, 0.05), 0.7, aspect)


def look_at_camera(eye, target, half, aspect):
    eye, target = np.asarray(eye, np.float64), np.asarray(target, np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    up = np.array([0.0, 1.0, 0.0])
    r = np.cross(up, f)
    r /= np.linalg.norm(r)
    u = np.cross(f, r)
    hor = 2 * half * aspect * r
    ver = 2 * half * u
    llc = eye + f - 0.5 * hor - 0.5 * ver
    return np.concatenate([eye, llc, hor, ver]).astype(np.float32)


def camera_rays(cam12, w, h):
    """(w*h, 6) float32 pixel-centre primary rays of a 12-float camera (origin, lower-left, horizontal, vertical), row 0 =
    bottom, the reference's GetRay formula (R/src/Camera.cu:18-20).  For the tools; the tests use the oracle's generator."""
    cam = np.asarray(cam12, np.float32)
    u = ((np.arange(w, dtype=np.float32) + np.float32(0.5)) / np.float32(w))[None, :, None]
    v = ((np.arange(h, dtype=np.float32) + np.float32(0.5)) / np.float32(h))[:, None, None]
    d = (cam[3:6] + u * cam[6:9] + v * cam[9:12] - cam[0:3]).astype(np.float32).reshape(-1, 3)
    o = np.broadcast_to(cam[0:3], d.shape)
    return np.ascontiguousarray(np.concatenate([o, d], 1), dtype=np.float32)

