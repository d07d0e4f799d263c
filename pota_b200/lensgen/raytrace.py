"""Sequential spherical-surface ray tracer used to author the lens pack (numpy, float64).

This replaces the role of polynomial-optics' `raytrace.h` (absent from /root/reference; it is
only referenced from dead code, /root/reference/src/deprecated/lentil_raytraced.cpp).  It is a
build-time tool: the GPU product and the oracle only ever see the fitted polynomials.

Coordinate system = the one the reference's polynomials live in (see how
/root/reference/src/lentil.h:381-389 feeds `out` to sphereToCs with centre -R): the optical axis
is z, light travels from the sensor towards +z, the vertex of the outermost (scene side) surface
is z = 0 and the scene is at z > 0.  Lengths in mm, wavelengths in micrometres.

Sensor-side rays are (x, y, dx, dy) in plane/plane parametrisation (direction (dx, dy, 1));
outer-pupil rays are in sphere/sphere parametrisation, i.e. what csToSphere
(/root/reference/src/lens.h:127-153) produces on the outer pupil sphere.
"""
from __future__ import annotations

import numpy as np

LAMBDA_D, LAMBDA_F, LAMBDA_C = 0.5876, 0.4861, 0.6563


def ior_at(n_d: float, abbe: float, lam: np.ndarray) -> np.ndarray:
    """Cauchy dispersion n(lambda) = A + B / lambda^2 from (n_d, V_d)."""
    lam = np.asarray(lam, dtype=np.float64)
    if n_d <= 1.0 or abbe <= 0.0:
        return np.ones_like(lam) * max(n_d, 1.0)
    b = (n_d - 1.0) / (abbe * (1.0 / LAMBDA_F**2 - 1.0 / LAMBDA_C**2))
    a = n_d - b / LAMBDA_D**2
    return a + b / lam**2


class Lens:
    """Scaled prescription + derived constants."""

    def __init__(self, surfaces, focal_mm: float | None = None):
        rows = [tuple(map(float, r)) for r in surfaces]
        self.rows = rows
        if focal_mm is not None:
            efl = Lens(rows).efl
            s = focal_mm / efl
            rows = [(r * s, d * s, n, v, h * s) for (r, d, n, v, h) in rows]
            self.rows = rows
        self.n = len(rows)
        self.R = np.array([r[0] for r in rows])
        self.d = np.array([r[1] for r in rows])
        self.nd = np.array([r[2] for r in rows])
        self.abbe = np.array([r[3] for r in rows])
        self.h = np.array([r[4] for r in rows])
        self.zv = -np.concatenate([[0.0], np.cumsum(self.d[:-1])])  # vertex z of each surface
        self.stop = int(np.where(self.R == 0.0)[0][0])
        self._paraxial()

    # -- paraxial constants -------------------------------------------------------------
    def _paraxial(self):
        eps = 1e-4
        lam = np.array([0.55])
        z_rear = self.zv[-1]
        # ray A: height eps at the rear vertex plane, parallel to the axis
        oa = self.trace_from(np.array([0.0]), np.array([eps]), np.array([0.0]), np.array([0.0]), lam, z_rear)
        # ray B: on axis at the rear vertex plane, slope eps
        ob = self.trace_from(np.array([0.0]), np.array([0.0]), np.array([0.0]), np.array([eps]), lam, z_rear)
        ua = oa["dir"][1, 0] / oa["dir"][2, 0] / eps  # exit slope per unit height
        ub = ob["dir"][1, 0] / ob["dir"][2, 0] / eps  # exit slope per unit slope
        # a ray leaving the axis at distance L behind the rear vertex with slope s has (y, u) = (sL, s)
        # at the rear vertex; it leaves parallel when ua*L + ub = 0
        self.bfl = float(-ub / ua)
        ya = (oa["pos"][1, 0] - oa["dir"][1, 0] / oa["dir"][2, 0] * oa["pos"][2, 0]) / eps
        yb = (ob["pos"][1, 0] - ob["dir"][1, 0] / ob["dir"][2, 0] * ob["pos"][2, 0]) / eps
        self.efl = float(abs(ya * self.bfl + yb))
        self.z_sensor = z_rear - self.bfl
        self.length = float(-self.z_sensor)

    # -- tracing ------------------------------------------------------------------------
    def trace_from(self, x, y, dx, dy, lam, z0, stop_at_aperture=False):
        """Trace rays starting on the plane z = z0 behind the lens towards +z.

        Returns dict(pos[3,N], dir[3,N], ok[N], T[N], ap[4,N]) with pos/dir on the outermost
        surface (after refraction) and ap = (x, y, dx, dy) on the aperture plane.
        """
        n_rays = x.shape[0]
        p = np.stack([x, y, np.full(n_rays, float(z0))]).astype(np.float64)
        dvec = np.stack([dx, dy, np.ones(n_rays)]).astype(np.float64)
        dvec /= np.linalg.norm(dvec, axis=0)
        ok = np.ones(n_rays, dtype=bool)
        T = np.ones(n_rays)
        ap = np.zeros((4, n_rays))
        with np.errstate(all="ignore"):
            for i in range(self.n - 1, -1, -1):
                R = self.R[i]
                zv = self.zv[i]
                if R == 0.0:  # aperture stop: plane
                    t = (zv - p[2]) / dvec[2]
                    p = p + dvec * t
                    ap[0], ap[1] = p[0], p[1]
                    ap[2], ap[3] = dvec[0] / dvec[2], dvec[1] / dvec[2]
                    ok &= (p[0] ** 2 + p[1] ** 2) <= self.h[i] ** 2
                    if stop_at_aperture:
                        break
                    continue
                c = zv - R  # sphere centre on the axis
                oc = p.copy()
                oc[2] -= c
                b = np.sum(oc * dvec, axis=0)
                cc = np.sum(oc * oc, axis=0) - R * R
                disc = b * b - cc
                ok &= disc >= 0.0
                sq = np.sqrt(np.maximum(disc, 0.0))
                # ray travels +z: vertex is the +z pole of the sphere when R > 0 (far root), else near root
                t = -b + sq if R > 0 else -b - sq
                p = p + dvec * t
                ok &= (p[0] ** 2 + p[1] ** 2) <= self.h[i] ** 2
                nrm = p.copy()
                nrm[2] -= c
                nrm /= R  # unit normal; points towards +z (along the ray) for either sign of R
                n1 = ior_at(self.nd[i], self.abbe[i], lam)  # medium behind surface i (we come from there)
                n2 = ior_at(self.nd[i - 1], self.abbe[i - 1], lam) if i > 0 else np.ones_like(lam)
                eta = n1 / n2
                cos1 = np.sum(dvec * nrm, axis=0)
                k = 1.0 - eta * eta * (1.0 - cos1 * cos1)
                ok &= k >= 0.0
                cos2 = np.sqrt(np.maximum(k, 0.0))
                dvec = eta * dvec + (cos2 - eta * cos1) * nrm
                dvec /= np.linalg.norm(dvec, axis=0)
                # unpolarised Fresnel transmittance
                rs = (n1 * cos1 - n2 * cos2) / (n1 * cos1 + n2 * cos2)
                rp = (n1 * cos2 - n2 * cos1) / (n1 * cos2 + n2 * cos1)
                T = T * (1.0 - 0.5 * (rs * rs + rp * rp))
        ok &= np.isfinite(p).all(axis=0) & np.isfinite(dvec).all(axis=0)
        return dict(pos=p, dir=dvec, ok=ok, T=T, ap=ap)

    def cs_to_sphere(self, pos, dvec):
        """csToSphere on the outer pupil (lens.h:127-153), centre -R0, radius R0."""
        R = self.R[0]
        c = -R
        nrm = np.stack([pos[0] / R, pos[1] / R, np.abs((pos[2] - c) / R)])
        ex = np.stack([nrm[2], np.zeros_like(nrm[0]), -nrm[0]])
        ex /= np.linalg.norm(ex, axis=0)
        ey = np.cross(nrm.T, ex.T).T
        d = dvec / np.linalg.norm(dvec, axis=0)
        return np.stack([pos[0], pos[1], np.sum(d * ex, axis=0), np.sum(d * ey, axis=0)])

    def sample(self, n_rays: int, sensor_half: float, seed: int = 1):
        """Random valid sensor rays (x,y,dx,dy,lambda) with their aperture / outer-pupil images."""
        rng = np.random.default_rng(seed)
        m = int(n_rays * 4)
        x = rng.uniform(-sensor_half, sensor_half, m)
        y = rng.uniform(-sensor_half, sensor_half, m)
        # aim at a uniformly sampled point on the rear element disc (slightly oversized)
        r = np.sqrt(rng.uniform(0, 1, m)) * self.h[-1] * 1.05
        phi = rng.uniform(0, 2 * np.pi, m)
        tx, ty = r * np.cos(phi), r * np.sin(phi)
        dx = (tx - x) / self.bfl
        dy = (ty - y) / self.bfl
        lam = rng.uniform(0.4, 0.7, m)
        o = self.trace_from(x, y, dx, dy, lam, self.z_sensor)
        ok = o["ok"]
        sph = self.cs_to_sphere(o["pos"], o["dir"])
        ok &= np.isfinite(sph).all(axis=0)
        idx = np.where(ok)[0][:n_rays]
        X = np.stack([x, y, dx, dy, lam])[:, idx]
        return X, o["ap"][:, idx], sph[:, idx], o["T"][idx]
