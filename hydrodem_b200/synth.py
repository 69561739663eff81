"""Synthetic SRTM / HydroSHEDS / groves rasters for parity tests and bench.py.

There is no network and the reference's input tiles are stripped from the
repo, so every workload above the bundled 519x508
tile is synthetic.  The recipe follows SURVEY.md section 8(d) with two
additions, stated in DESIGN.md: a 0.3 m white-noise term on the SRTM surface
(real SRTM carries ~1 m of sensor noise) and 0.55 m of texture on the
HydroSHEDS surface before rounding (calibrated on the bundled tile), with
twice the lagoon density so that ~7 % of the cells are lagoon plateaus.  Without it the high-frequency half
of the spectrum sits at the float32 FFT rounding floor and the Fourier
peak detector (centre > 4 x hollow mean) thresholds pure rounding noise.

Everything is a pure function of (ny, nx, seed) and can be produced one row
band at a time (``rows=(r0, r1)``), which is how the row-band sharded runs
and the 36000^2 mosaic generate their inputs.
"""
import numpy as np

_BLOCK = 256          # row granularity of the per-cell random streams


def _bilinear(coarse, step, r0, r1, nx):
    """Rows r0..r1 of ``coarse`` (nodes every ``step`` cells) upsampled bilinearly."""
    y = np.arange(r0, r1, dtype=np.float64) / step
    x = np.arange(nx, dtype=np.float64) / step
    iy = y.astype(np.int64); fy = (y - iy).astype(np.float32)
    ix = x.astype(np.int64); fx = (x - ix).astype(np.float32)
    top = coarse[iy][:, ix] * (1 - fx) + coarse[iy][:, ix + 1] * fx
    bot = coarse[iy + 1][:, ix] * (1 - fx) + coarse[iy + 1][:, ix + 1] * fx
    return top * (1 - fy)[:, None] + bot * fy[:, None]


class SynthScene:
    """One synthetic scene: SRTM surface with stripes and groves, the matching
    HydroSHEDS raster with lagoon plateaus and voids, the groves class and an
    (empty) rivers raster."""

    def __init__(self, ny, nx, seed):
        self.ny, self.nx, self.seed = int(ny), int(nx), int(seed)
        rng = np.random.default_rng([self.seed, 1])
        self._octaves = []
        for k in range(6):
            step = 2 ** (9 - k)
            cy, cx = self.ny // step + 3, self.nx // step + 3
            self._octaves.append((step, np.float32(8.0 * 0.5 ** k),
                                  rng.standard_normal((cy, cx)).astype(np.float32)))
        cells = self.ny * self.nx
        # groves: thin rectangles, ~1 % coverage
        rg = np.random.default_rng([self.seed, 2])
        n_groves = max(1, cells // 40000)
        self._groves = []
        for _ in range(n_groves):
            hgt, wid = (int(rg.integers(3, 9)), int(rg.integers(20, 120)))
            if rg.random() < 0.5:
                hgt, wid = wid, hgt
            hgt, wid = min(hgt, self.ny), min(wid, self.nx)
            y0 = int(rg.integers(0, self.ny - hgt + 1)); x0 = int(rg.integers(0, self.nx - wid + 1))
            self._groves.append((y0, x0, hgt, wid, np.float32(rg.uniform(2.0, 6.0))))
        # lagoons: discs flattened to their minimum
        rl = np.random.default_rng([self.seed, 3])
        n_lag = max(1, cells // 25000)
        self._lagoons = [(int(rl.integers(0, self.ny)), int(rl.integers(0, self.nx)),
                          float(rl.uniform(6.0, 40.0))) for _ in range(n_lag)]
        self._lagoon_level = None
        # voids: -32768 at density 2e-5, never on the frame
        rv = np.random.default_rng([self.seed, 4])
        n_void = max(1, int(round(cells * 2e-5))) if min(self.ny, self.nx) > 2 else 0
        self._voids = np.stack([rv.integers(1, max(2, self.ny - 1), n_void),
                                rv.integers(1, max(2, self.nx - 1), n_void)], axis=1)

    # ---- pieces -----------------------------------------------------------------
    def _rows(self, rows):
        r0, r1 = (0, self.ny) if rows is None else rows
        assert 0 <= r0 <= r1 <= self.ny
        return r0, r1

    def base(self, rows=None):
        """Smooth terrain: trend + 6 octaves of bilinear value noise (float32)."""
        r0, r1 = self._rows(rows)
        x = np.arange(self.nx, dtype=np.float32)
        y = np.arange(r0, r1, dtype=np.float32)
        out = (np.float32(100.0) + np.float32(1e-4) * x[None, :] + np.float32(5e-5) * y[:, None]).astype(np.float32)
        for step, amp, coarse in self._octaves:
            out += amp * _bilinear(coarse, step, r0, r1, self.nx).astype(np.float32)
        return out

    def _white(self, rows, stream, scale):
        r0, r1 = self._rows(rows)
        out = np.empty((r1 - r0, self.nx), dtype=np.float32)
        b = r0 // _BLOCK
        while b * _BLOCK < r1:
            rng = np.random.default_rng([self.seed, stream, b])
            blk = rng.standard_normal((_BLOCK, self.nx), dtype=np.float32)
            lo, hi = max(r0, b * _BLOCK), min(r1, (b + 1) * _BLOCK)
            out[lo - r0:hi - r0] = blk[lo - b * _BLOCK:hi - b * _BLOCK]
            b += 1
        return out * np.float32(scale)

    def groves(self, rows=None):
        """Groves classification, uint8 0/1 (image_srtm.py:177 reads it with GDAL)."""
        r0, r1 = self._rows(rows)
        g = np.zeros((r1 - r0, self.nx), dtype=np.uint8)
        for (y0, x0, hgt, wid, _) in self._groves:
            a, b = max(y0, r0), min(y0 + hgt, r1)
            if a < b:
                g[a - r0:b - r0, x0:x0 + wid] = 1
        return g

    def srtm(self, rows=None):
        """Raw SRTM: base + stripes + sensor noise + tree canopy over groves (float32)."""
        r0, r1 = self._rows(rows)
        out = self.base(rows)
        x = np.arange(self.nx, dtype=np.float64)[None, :]
        y = np.arange(r0, r1, dtype=np.float64)[:, None]
        out += (0.5 * np.sin(2 * np.pi * (0.11 * x + 0.07 * y))
                + 0.3 * np.sin(2 * np.pi * (0.031 * x - 0.052 * y))).astype(np.float32)
        out += self._white(rows, 5, 0.3)
        for (y0, x0, hgt, wid, canopy) in self._groves:
            a, b = max(y0, r0), min(y0 + hgt, r1)
            if a < b:
                out[a - r0:b - r0, x0:x0 + wid] += canopy
        return out

    def _levels(self):
        if self._lagoon_level is None:
            lev = []
            for (cy, cx, rad) in self._lagoons:
                r = int(np.ceil(rad))
                a, b = max(cy - r, 0), min(cy + r + 1, self.ny)
                blk = np.round(self.base((a, b)))
                yy = np.arange(a, b)[:, None] - cy
                xx = np.arange(self.nx)[None, :] - cx
                disc = (yy * yy + xx * xx) <= rad * rad
                lev.append(np.float32(blk[disc].min()) if disc.any() else np.float32(0))
            self._lagoon_level = lev
        return self._lagoon_level

    def hsheds(self, rows=None):
        """HydroSHEDS-style raster: integer metres as float32, lagoon plateaus, voids."""
        r0, r1 = self._rows(rows)
        # 0.55 m of texture before rounding to integer metres: the bundled HydroSHEDS tile has ~59 % equal
        # horizontal neighbours (a perfectly smooth surface would make MajorityFilter fire everywhere)
        out = np.round(self.base(rows) + self._white(rows, 6, 0.55)).astype(np.float32)
        for (cy, cx, rad), lev in zip(self._lagoons, self._levels()):
            r = int(np.ceil(rad))
            a, b = max(cy - r, r0), min(cy + r + 1, r1)
            if a >= b:
                continue
            x0, x1 = max(cx - r, 0), min(cx + r + 1, self.nx)
            yy = np.arange(a, b)[:, None] - cy
            xx = np.arange(x0, x1)[None, :] - cx
            disc = (yy * yy + xx * xx) <= rad * rad
            blk = out[a - r0:b - r0, x0:x1]
            blk[disc] = lev
        for (vy, vx) in self._voids:
            if r0 <= vy < r1:
                out[vy - r0, vx] = np.float32(-32768.0)
        return out

    def rivers(self, rows=None):
        """Rasterised rivers: zeros (RouteRivers is out of scope)."""
        r0, r1 = self._rows(rows)
        return np.zeros((r1 - r0, self.nx), dtype=np.float32)


def make_scene(ny, nx, seed):
    return SynthScene(ny, nx, seed)


# seeds fixed by SURVEY.md section 8(d)
CONFIG_SEEDS = {"C2": 1002, "C3": 1003, "C4": 1004, "C5": 1005}


# ---- the same recipe generated ON THE DEVICE, any row band at a time -------------------------------------------------
class DeviceMosaic:
    """Synthetic mosaic of the BASELINE.json shapes (18000^2, 36000^2) generated with torch on the current CUDA device.

    The NumPy generator above needs ~25 minutes for 36000 x 36000 cells; this one needs seconds.  Same recipe (trend +
    6 octaves of bilinear value noise + stripes + sensor noise + canopy; HydroSHEDS = rounded textured base with lagoon
    plateaus and voids), NOT the same random numbers.  Every cell is a pure function of (y, x, seed): the per-cell
    noise comes from an integer hash of the coordinates, lattices / rectangles / discs from seeded host generators,
    so ``band(r0, r1)`` returns the same values whichever rank asks and however the mosaic is cut -- the row-band
    sharded bench runs at N = 1, 2, 4, 8 all work on ONE mosaic (strong scaling, identical checksums)."""

    def __init__(self, ny, nx, seed, device=None):
        import torch
        self.ny, self.nx, self.seed = int(ny), int(nx), int(seed)
        self.torch = torch
        dev_ = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev_
        rng = np.random.default_rng([self.seed, 11])
        self.octaves = []
        for k in range(6):
            step = 2 ** (9 - k)
            cy, cx = self.ny // step + 3, self.nx // step + 3
            lat = torch.from_numpy(rng.standard_normal((cy, cx)).astype(np.float32)).to(dev_)
            self.octaves.append((step, 8.0 * 0.5 ** k, lat))
        cells = self.ny * self.nx
        rg = np.random.default_rng([self.seed, 12])
        n = max(1, cells // 40000)
        hgt, wid = rg.integers(3, 9, n), rg.integers(20, 120, n)
        swap = rg.random(n) < 0.5
        hgt, wid = np.where(swap, wid, hgt), np.where(swap, hgt, wid)
        hgt, wid = np.minimum(hgt, self.ny), np.minimum(wid, self.nx)
        self.g_y0 = (rg.random(n) * (self.ny - hgt + 1)).astype(np.int64)
        self.g_x0 = (rg.random(n) * (self.nx - wid + 1)).astype(np.int64)
        self.g_h, self.g_w, self.g_canopy = hgt.astype(np.int64), wid.astype(np.int64), rg.uniform(2.0, 6.0, n).astype(np.float32)
        rl = np.random.default_rng([self.seed, 13])
        n = max(1, cells // 25000)
        self.l_cy, self.l_cx = rl.integers(41, max(42, self.ny - 41), n), rl.integers(41, max(42, self.nx - 41), n)
        self.l_rad = rl.uniform(6.0, 40.0, n)
        rv = np.random.default_rng([self.seed, 14])
        n = max(1, int(round(cells * 2e-5)))
        self.v_y, self.v_x = rv.integers(1, self.ny - 1, n), rv.integers(1, self.nx - 1, n)

    def _normal(self, ys, xs, stream):
        """Standard normal per cell from an integer hash of (y, x, seed, stream): Box-Muller on two 24-bit uniforms."""
        torch = self.torch
        m = 0xFFFFFFFF

        def mix(h):
            h = h ^ (h >> 15)
            h = (h * 0x2C1B3C6D) & m
            h = h ^ (h >> 12)
            h = (h * 0x297A2D39) & m
            return h ^ (h >> 15)

        key = (self.seed * 0x9E3779B1 + stream * 0x7F4A7C15) & m
        h0 = (ys[:, None] * 0x85EBCA77 + xs[None, :] * 0xC2B2AE3D + key) & m
        a, b = mix(h0), mix(h0 ^ 0x68E31DA4)
        u1 = ((a >> 8).to(torch.float32) + 1.0) * (1.0 / 16777217.0)
        u2 = (b >> 8).to(torch.float32) * (1.0 / 16777216.0)
        return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)

    def _base(self, ys, xs):
        torch = self.torch
        import torch.nn.functional as F
        yf, xf = ys.to(torch.float64), xs.to(torch.float64)
        base = (100.0 + 1e-4 * xf[None, :] + 5e-5 * yf[:, None]).float()
        for step, amp, lat in self.octaves:
            gy = (yf / step / (lat.shape[0] - 1) * 2 - 1).float()
            gx = (xf / step / (lat.shape[1] - 1) * 2 - 1).float()
            grid = torch.stack(torch.broadcast_tensors(gx[None, :], gy[:, None]), dim=-1)[None]
            base += amp * F.grid_sample(lat[None, None], grid, mode="bilinear", align_corners=True)[0, 0]
        return base

    def _fill(self, r0, r1, srtm, hsheds, chunk=1024):
        torch = self.torch
        xs = torch.arange(self.nx, device=self.device, dtype=torch.int64)
        for a in range(r0, r1, chunk):
            b = min(r1, a + chunk)
            ys = torch.arange(a, b, device=self.device, dtype=torch.int64)
            base = self._base(ys, xs)
            if srtm is not None:
                yf, xf = ys.to(torch.float64), xs.to(torch.float64)
                stripes = (0.5 * torch.sin(2 * np.pi * (0.11 * xf[None, :] + 0.07 * yf[:, None]))
                           + 0.3 * torch.sin(2 * np.pi * (0.031 * xf[None, :] - 0.052 * yf[:, None]))).float()
                srtm[a - r0:b - r0] = base + stripes + 0.3 * self._normal(ys, xs, 5)
            if hsheds is not None:
                hsheds[a - r0:b - r0] = torch.round(base + 0.55 * self._normal(ys, xs, 6))

    def band(self, r0, r1, out=None):
        """(srtm F32, groves U8, hsheds F32) torch tensors of rows [r0, r1).  ``out``: optional dict of preallocated
        (rows x nx) tensor views to fill instead (keys srtm / groves / hsheds)."""
        torch = self.torch
        rows = r1 - r0
        out = out or {}
        srtm = out.get("srtm")
        groves = out.get("groves")
        hsheds = out.get("hsheds")
        if srtm is None:
            srtm = torch.empty((rows, self.nx), dtype=torch.float32, device=self.device)
        if groves is None:
            groves = torch.empty((rows, self.nx), dtype=torch.uint8, device=self.device)
        if hsheds is None:
            hsheds = torch.empty((rows, self.nx), dtype=torch.float32, device=self.device)
        self._fill(r0, r1, srtm, hsheds)
        groves.zero_()
        sel = np.nonzero((self.g_y0 < r1) & (self.g_y0 + self.g_h > r0))[0]
        for k in sel:
            a, b = max(int(self.g_y0[k]), r0) - r0, min(int(self.g_y0[k] + self.g_h[k]), r1) - r0
            x0, x1 = int(self.g_x0[k]), int(self.g_x0[k] + self.g_w[k])
            groves[a:b, x0:x1] = 1
            srtm[a:b, x0:x1] += float(self.g_canopy[k])
        # lagoon plateaus: level = (unmodified) rounded texture at the disc centre - 2, so it does not depend on which
        # other discs a rank happens to see
        sel = np.nonzero((self.l_cy + 41 > r0) & (self.l_cy - 41 < r1))[0]
        if len(sel):
            cy = torch.from_numpy(self.l_cy[sel]).to(self.device)
            cx = torch.from_numpy(self.l_cx[sel]).to(self.device)
            lev = (self._point_texture(cy, cx) - 2.0).cpu().numpy()      # centre values, evaluated point-wise
            yy, xx = torch.meshgrid(torch.arange(-40, 41, device=self.device), torch.arange(-40, 41, device=self.device),
                                    indexing="ij")
            rr = (yy * yy + xx * xx).float()
            for i, k in enumerate(sel):
                cyk, cxk, rad = int(self.l_cy[k]), int(self.l_cx[k]), float(self.l_rad[k])
                a, b = max(cyk - 40, r0), min(cyk + 41, r1)
                if a >= b:
                    continue
                blk = hsheds[a - r0:b - r0, cxk - 40:cxk + 41]
                disc = rr[a - (cyk - 40):b - (cyk - 40)] <= rad * rad
                blk[disc] = float(lev[i])
        vs = np.nonzero((self.v_y >= r0) & (self.v_y < r1))[0]
        if len(vs):
            hsheds[torch.from_numpy(self.v_y[vs] - r0).to(self.device), torch.from_numpy(self.v_x[vs]).to(self.device)] = -32768.0
        return srtm, groves, hsheds

    def _point_texture(self, ys, xs):
        """round(base + 0.55 * noise) at scattered points (ys[i], xs[i]), vectorised."""
        torch = self.torch
        import torch.nn.functional as F
        yf, xf = ys.to(torch.float64), xs.to(torch.float64)
        base = (100.0 + 1e-4 * xf + 5e-5 * yf).float()
        for step, amp, lat in self.octaves:
            gy = (yf / step / (lat.shape[0] - 1) * 2 - 1).float()
            gx = (xf / step / (lat.shape[1] - 1) * 2 - 1).float()
            grid = torch.stack((gx, gy), dim=-1)[None, None]
            base += amp * F.grid_sample(lat[None, None], grid, mode="bilinear", align_corners=True)[0, 0, 0]
        m = 0xFFFFFFFF

        def mix(h):
            h = h ^ (h >> 15)
            h = (h * 0x2C1B3C6D) & m
            h = h ^ (h >> 12)
            h = (h * 0x297A2D39) & m
            return h ^ (h >> 15)

        key = (self.seed * 0x9E3779B1 + 6 * 0x7F4A7C15) & m
        h0 = (ys * 0x85EBCA77 + xs * 0xC2B2AE3D + key) & m
        a, b = mix(h0), mix(h0 ^ 0x68E31DA4)
        u1 = ((a >> 8).to(torch.float32) + 1.0) * (1.0 / 16777217.0)
        u2 = (b >> 8).to(torch.float32) * (1.0 / 16777216.0)
        return torch.round(base + 0.55 * torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2))
