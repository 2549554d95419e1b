"""Host-side world model: Python mirror of the reference's World grid, world generators and save files.

  World                     Core/World.h:35-70,192 (m_WorldData, GetBlock/SetBlock), x-fastest layout; CPU picking
                            (World::Raycast / RaycastDetect, Core/World.cpp:215-546)
  generate_superflat/plains Core/WorldGenerator.cpp:28-121
  save_world / load_world   Core/WorldFileHandler.cpp:10-83 (raw 18,874,368-byte dump)
  generate_gi_box / city    deterministic stand-ins for the two scenes whose blobs are missing from the reference
                            checkout
"""
import os

import numpy as np

from .abi import WORLD_SIZE_X, WORLD_SIZE_Y, WORLD_SIZE_Z, WORLD_VOXELS

# block ids = position in blockdb.txt (Core/BlockDatabaseParser.cpp:31-42)
GRASS, DIRT, STONE, COBBLESTONE, SAND, LEAVES, LAMP, PLANKS, GLOWSTONE = 1, 2, 3, 4, 5, 7, 12, 20, 27


class World:
    """384 x 128 x 384 uint8 block ids; idx = x + 384*y + 49152*z (Core/World.h:46-48). 0 = air."""

    def __init__(self, data=None):
        if data is None:
            self.data = np.zeros(WORLD_VOXELS, dtype=np.uint8)
        else:
            data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
            if data.size != WORLD_VOXELS:
                raise ValueError(f"world must hold {WORLD_VOXELS} bytes, got {data.size}")
            self.data = data

    @property
    def zyx(self):
        """[z][y][x] view of the same memory."""
        return self.data.reshape(WORLD_SIZE_Z, WORLD_SIZE_Y, WORLD_SIZE_X)

    def get_block(self, x, y, z):
        return int(self.data[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y])

    def set_block(self, x, y, z, block):
        if not (0 <= x < WORLD_SIZE_X and 0 <= y < WORLD_SIZE_Y and 0 <= z < WORLD_SIZE_Z):
            raise IndexError("voxel outside the world")
        self.data[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y] = block

    # reference spellings
    GetBlock = get_block
    SetBlock = set_block

    # ---- CPU picking: World::RaycastDetect / World::Raycast (Core/World.cpp:215-546) --------------------------------------------
    # The block the player looks at, found on the HOST grid by stepping from cell face to cell face (reach: 48 steps); it does not read
    # the distance field.  fp32 arithmetic in the reference's order (glm::vec3); std::min(a, b) = (b < a) ? b : a, so a NaN quotient
    # (0 / 0 on an axis the ray does not move along) loses against any number that follows it exactly as in the reference.
    @staticmethod
    def _pick_steps(pos, direction):
        f32 = np.float32
        p = [f32(v) for v in pos]
        d = [f32(v) for v in direction]
        sign = [f32(1.0) if c > 0 else f32(0.0) for c in d]

        def smin(a, b):
            return b if b < a else a

        with np.errstate(divide="ignore", invalid="ignore"):
            for _ in range(48):
                tvec = [(np.floor(p[k] + sign[k]) - p[k]) / d[k] for k in range(3)]
                t = smin(tvec[0], smin(tvec[1], tvec[2]))
                p = [p[k] + d[k] * (t + f32(0.001)) for k in range(3)]
                yield p, t, tvec, sign

    @staticmethod
    def _cell(p):
        """(int)floor(position) per axis and the reference's bounds verdict (>= size or <= 0 is outside: cell 0 counts as outside)."""
        c = []
        for v in p:
            c.append(int(np.floor(v)) if np.isfinite(v) else -2**31)
        inside = not (c[0] >= WORLD_SIZE_X or c[1] >= WORLD_SIZE_Y or c[2] >= WORLD_SIZE_Z or c[0] <= 0 or c[1] <= 0 or c[2] <= 0)
        return c, inside

    def raycast_detect(self, pos, direction):
        """RaycastDetect (:497-546): (x, y, z, block) of the first solid cell along the ray, (-1, -1, -1, -1) when the walk ends outside
        the world at a hit, None when nothing is hit within reach (the reference falls off the end of the function there)."""
        for p, _, _, _ in self._pick_steps(pos, direction):
            c, inside = self._cell(p)
            if inside and self.get_block(*c) != 0:
                return (c[0], c[1], c[2], self.get_block(*c))
        return None

    def raycast(self, op, pos, direction, held_block=STONE):
        """World::Raycast (:215-495): op 0 breaks the block looked at, 1 places `held_block` on the face looked at (unless it would
        intersect the player standing at `pos`), 2 picks the block looked at.  Edits the HOST grid only and returns
        {"changed": bool (the reference's return value), "voxel": (x, y, z) or None, "block": edited / picked id}; the caller mirrors
        an edit to the device with Renderer.set_block + build_distance_field, as the reference does with glTexSubImage3D +
        GenerateDistanceField (:372-374, 458-460)."""
        f32 = np.float32
        origin = [f32(v) for v in pos]
        for p, t, tvec, sign in self._pick_steps(pos, direction):
            c, inside = self._cell(p)
            if not (inside and self.get_block(*c) != 0):
                continue
            normal = [f32(1.0) if t == tvec[k] else f32(0.0) for k in range(3)]
            normal = [-normal[k] if sign[k] else normal[k] for k in range(3)]
            if op == 1:
                p = [p[k] + normal[k] for k in range(3)]
            p = [np.floor(v) for v in p]
            c, inside = self._cell(p)
            if not inside:
                return {"changed": False, "voxel": None, "block": 0}
            if op == 1:
                def dist(a, b):
                    dx, dy, dz = (f32(b[k]) - f32(a[k]) for k in range(3))
                    return np.sqrt(f32(f32(dx * dx + dy * dy) + dz * dz))

                ox, oy, oz = int(origin[0]), int(origin[1]), int(origin[2])
                under = [self.get_block(ox, int(f32(oy) - f32(1.0)), oz), self.get_block(ox, int(origin[1] - f32(1.0)), oz),
                         self.get_block(ox, int(origin[1] - f32(1.1)), oz)]
                if self.get_block(*c) != 0 or dist(p, origin) < f32(1.25):
                    return {"changed": False, "voxel": None, "block": 0}
                if 0 in under and dist(p, [origin[0], origin[1] - f32(1.0), origin[2]]) < f32(1.35):
                    return {"changed": False, "voxel": None, "block": 0}
                changed = self.get_block(*c) != held_block
                self.set_block(c[0], c[1], c[2], held_block)
                return {"changed": bool(changed), "voxel": tuple(c), "block": int(held_block)}
            if op == 0:
                old = self.get_block(*c)
                self.set_block(c[0], c[1], c[2], 0)
                return {"changed": True, "voxel": tuple(c), "block": old}
            if op == 2:
                return {"changed": False, "voxel": tuple(c), "block": self.get_block(*c)}
            return {"changed": False, "voxel": None, "block": 0}
        return {"changed": False, "voxel": None, "block": 0}


def _set_vertical_blocks(zyx, x, z, y_level, biome):
    """SetVerticalBlocks, Core/WorldGenerator.cpp:28-67."""
    y_level = min(int(y_level), WORLD_SIZE_Y)
    if y_level <= 0:
        return
    col = zyx[z, :y_level, x]
    col[:] = STONE
    if biome == 1:
        col[max(y_level - 5, 0):] = DIRT
        col[max(y_level - 1, 0):] = GRASS
    else:
        col[max(y_level - 8, 0):] = SAND


def generate_superflat():
    """GenerateWorld(world, false): every column 50 high, grass / 4 dirt / stone (WorldGenerator.cpp:109-121)."""
    w = World()
    v = w.zyx
    v[:, 0:45, :] = STONE
    v[:, 45:49, :] = DIRT
    v[:, 49, :] = GRASS
    return w


def generate_plains(columns):
    """GenerateWorld(world, true) from the per-column (height, biome) table the reference's FastNoise produces
    (voxelpathtracer_b200/data/plains_columns.u8, see tools/make_fixtures.py; WorldGenerator.cpp:85-107)."""
    cols = np.asarray(columns, dtype=np.uint8).reshape(WORLD_SIZE_X, WORLD_SIZE_Z, 2)
    w = World()
    v = w.zyx
    y = np.arange(WORLD_SIZE_Y, dtype=np.int32)[None, :, None]           # [1][y][1]
    h = cols[:, :, 0].astype(np.int32).T[:, None, :]                     # [z][1][x]
    grass = (cols[:, :, 1].T == 1)[:, None, :]
    solid = y < h
    blocks = np.where(y >= h - 1, GRASS, np.where(y >= h - 5, DIRT, STONE))
    sand = np.where(y >= h - 8, SAND, STONE)
    v[...] = np.where(solid, np.where(grass, blocks, sand), 0).astype(np.uint8)
    return w


def _lcg(seed):
    state = seed & 0xFFFFFFFF
    while True:
        state = (1664525 * state + 1013904223) & 0xFFFFFFFF
        yield state >> 8


def generate_gi_box(columns, seed=1234, rooms=40):
    """Stand-in for the missing 'Test Worlds/gi' save: plains + hollow stone rooms with a doorway and an
    emissive Lamp block on the ceiling, placed by a fixed LCG."""
    w = generate_plains(columns)
    v = w.zyx
    rnd = _lcg(seed)
    for _ in range(rooms):
        sx = 8 + next(rnd) % 24
        sy = 6 + next(rnd) % 10
        sz = 8 + next(rnd) % 24
        x0 = 4 + next(rnd) % (WORLD_SIZE_X - sx - 8)
        z0 = 4 + next(rnd) % (WORLD_SIZE_Z - sz - 8)
        y0 = 56 + next(rnd) % 8
        v[z0:z0 + sz, y0:y0 + sy, x0:x0 + sx] = COBBLESTONE
        v[z0 + 1:z0 + sz - 1, y0 + 1:y0 + sy - 1, x0 + 1:x0 + sx - 1] = 0
        v[z0 + sz // 2 - 1:z0 + sz // 2 + 1, y0 + 1:y0 + 4, x0] = 0                      # doorway on -x
        v[z0 + sz // 2, y0 + sy - 2, x0 + sx // 2] = LAMP                                # ceiling lamp
        v[z0:z0 + sz, 40:y0, x0:x0 + sx][v[z0:z0 + sz, 40:y0, x0:x0 + sx] == 0] = STONE  # foundation
    return w


def generate_orchard(columns, seed=77, trees=260):
    """Plains with trees (plank trunks, canopies of the Transparent leaves block) placed by a fixed LCG: the scene of the
    alpha-tested traversal tests (u_ShouldAlphaTest: rays pass the cut-out texels of Transparent blocks)."""
    w = generate_plains(columns)
    v = w.zyx
    cols = np.asarray(columns, dtype=np.uint8).reshape(WORLD_SIZE_X, WORLD_SIZE_Z, 2)
    rnd = _lcg(seed)
    for _ in range(trees):
        x = 8 + next(rnd) % (WORLD_SIZE_X - 16)
        z = 8 + next(rnd) % (WORLD_SIZE_Z - 16)
        h = int(cols[x, z, 0])
        trunk = 4 + next(rnd) % 4
        r = 2 + next(rnd) % 2
        top = h + trunk
        canopy = v[z - r:z + r + 1, top - 2:top + 2, x - r:x + r + 1]
        canopy[canopy == 0] = LEAVES
        v[z - 1:z + 2, top + 2, x - 1:x + 2][v[z - 1:z + 2, top + 2, x - 1:x + 2] == 0] = LEAVES
        v[z, h:top, x] = PLANKS
    return w


def generate_city(seed=99):
    """Stand-in for the missing 'Medival' Minecraft import: a dense deterministic grid of towers, arches and
    interiors up to y = 120 (fill ratio >= 25 %)."""
    w = World()
    v = w.zyx
    v[:, 0:40, :] = STONE
    v[:, 40, :] = COBBLESTONE
    rnd = _lcg(seed)
    pitch = 24
    for bz in range(0, WORLD_SIZE_Z, pitch):
        for bx in range(0, WORLD_SIZE_X, pitch):
            fx = 14 + next(rnd) % 7
            fz = 14 + next(rnd) % 7
            hgt = 20 + next(rnd) % 60
            mat = (STONE, COBBLESTONE, PLANKS, SAND)[next(rnd) % 4]
            x0, z0, y0 = bx + 2, bz + 2, 41
            x1, z1, y1 = min(x0 + fx, WORLD_SIZE_X), min(z0 + fz, WORLD_SIZE_Z), min(y0 + hgt, 120)
            v[z0:z1, y0:y1, x0:x1] = mat
            for fy in range(y0 + 1, y1 - 1, 6):  # hollow floors with windows and a lamp
                v[z0 + 1:z1 - 1, fy:fy + 5, x0 + 1:x1 - 1] = 0
                v[z0 + 2:z1 - 2:3, fy + 1:fy + 3, x0] = 0
                v[z0, fy + 1:fy + 3, x0 + 2:x1 - 2:3] = 0
                if (fy // 6) % 2 == 0 and z1 - z0 > 4 and x1 - x0 > 4:
                    v[(z0 + z1) // 2, fy + 4, (x0 + x1) // 2] = LAMP
            if next(rnd) % 2 == 0 and x1 + 6 < WORLD_SIZE_X:  # arch to the next lot
                ay = min(y0 + 10 + next(rnd) % 20, 118)
                v[z0 + 2:z0 + 5, ay:ay + 2, x1:min(x1 + pitch - fx, WORLD_SIZE_X)] = COBBLESTONE
    return w


def save_world(world, path):
    """SaveWorld (Core/WorldFileHandler.cpp:10-38): the raw grid bytes."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    world.data.tofile(path)
    return True


def load_world(path):
    """LoadWorld (Core/WorldFileHandler.cpp:40-83) without the light-list scan (out of scope)."""
    data = np.fromfile(path, dtype=np.uint8)
    if data.size != WORLD_VOXELS:
        raise ValueError(f"{path}: expected {WORLD_VOXELS} bytes, found {data.size}")
    return World(data)
