"""Host-side scalar time-step controller: ``timestep`` and ``expansion`` of CUBE/main/timestep.f90:1-133 and the
initial values of CUBE/main/initialize.f90:25-37.  Scalars only -- this stays on the host; it
consumes the four limits ``particle_mesh`` returns and produces ``dt_old, dt, a_mid`` for the next step.

Fortran default real is f32: every intermediate that the reference keeps in a ``real`` variable is rounded to f32 here;
``expansion`` works in real(8) between its f32 arguments and results (timestep.f90:93-94).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f64 = np.float64


class Cosmology:
    """parameters.f90:63-86."""

    def __init__(self, z_i=49.0, omega_c=0.27, omega_b=0.05, wde=-1.0, ra_max=0.2, dt_max=1.0, box=200.0, h0=67.0):
        self.z_i, self.wde, self.ra_max, self.dt_max, self.box, self.h0 = f32(z_i), f32(wde), f32(ra_max), f32(dt_max), f32(box), f32(h0)
        self.omega_m = f32(omega_c) + f32(omega_b)
        self.omega_l = f32(1) - self.omega_m


def expansion(cos: Cosmology, a0, dt0):
    """timestep.f90:89-133: third-order Taylor step of the Friedmann equation, two half steps."""
    a0, dt0 = f32(a0), f32(dt0)
    dt_x = f32(dt0 / f32(2))
    dt2, dt3 = f32(dt_x * dt_x), f32(f32(dt_x * dt_x) * dt_x)      # dt_x**2, dt_x**3: real(4) powers
    om, ol, w = cos.omega_m, cos.omega_l, cos.wde
    omHsq = f64(f32(4.0) / f32(9.0))
    c_add = f64(f32(1.5) * (f32(1.0) - w))
    c_atd = f64(f32(1.5) * (f32(2.0) - f32(3.0) * w) * (f32(1.0) - w))

    def half(a_x):
        a3rlm = a_x ** f64(f32(-3) * w) * f64(ol) / f64(om)
        arkm = a_x * f64(f32(1.0) - om - ol) / f64(om)
        adot = np.sqrt(omHsq * (a_x * a_x * a_x) * (1.0 + arkm + a3rlm))
        addot = (a_x * a_x) * omHsq * (1.5 + 2.0 * arkm + c_add * a3rlm)
        atdot = a_x * adot * omHsq * (3.0 + 6.0 * arkm + c_atd * a3rlm)
        return f32(adot * f64(dt_x) + (addot * f64(dt2)) / 2.0 + (atdot * f64(dt3)) / 6.0)

    da1 = half(f64(a0))
    da2 = half(f64(f32(a0 + da1)))
    return da1, da2


class TimeStepper:
    """State of timestep.f90 between calls; :meth:`step` is one ``call timestep``."""

    def __init__(self, cos: Cosmology, z_checkpoint):
        self.cos = cos
        self.z_checkpoint = [f32(z) for z in z_checkpoint]
        self.a = f32(1) / (f32(1) + cos.z_i)
        self.a_mid = self.a
        self.dt = self.dt_old = self.da = self.t = f32(0)
        self.tau = f32(-3) / np.sqrt(self.a)
        self.dt_fine = self.dt_coarse = self.dt_pp = self.dt_vmax = f32(1000)
        self.cur_checkpoint, self.checkpoint_step, self.final_step, self.istep = 0, False, False, 0

    def limits(self, pm):
        """Take the dt limits a ``particle_mesh`` call returned (pm.f90:233-244)."""
        self.dt_fine, self.dt_coarse, self.dt_vmax = f32(pm["dt_fine"]), f32(pm["dt_coarse"]), f32(pm["dt_vmax"])

    def step(self):
        c = self.cos
        self.dt_old = self.dt
        dt_e, ntemp = c.dt_max, 0
        while True:                                            # timestep.f90:17-29
            ntemp += 1
            da1, da2 = expansion(c, self.a, dt_e)
            da = f32(da1 + da2)
            ra = f32(da / f32(self.a + da))
            if ra > c.ra_max:
                dt_e = f32(dt_e * f32(c.ra_max / ra))
            else:
                break
            if ntemp > 10:
                break
        dt = min(dt_e, self.dt_fine, self.dt_coarse, self.dt_pp, self.dt_vmax)
        da1, da2 = expansion(c, self.a, dt)
        da = f32(da1 + da2)
        self.checkpoint_step = False
        a_chk = f32(1.0) / f32(f32(1) + self.z_checkpoint[self.cur_checkpoint])
        if da >= f32(a_chk - self.a):                          # timestep.f90:49-58
            self.checkpoint_step = True
            if self.cur_checkpoint == len(self.z_checkpoint) - 1:
                self.final_step = True
            for _ in range(100):
                if abs(f32(f32(self.a + da) / a_chk) - f32(1)) < f32(1e-6):
                    break
                dt = f32(f32(dt * f32(a_chk - self.a)) / da)
                da1, da2 = expansion(c, self.a, dt)
                da = f32(da1 + da2)
            else:   # the reference's `do while` has no cap (timestep.f90:53-58): it would spin; say so instead of stepping past a_checkpoint
                raise RuntimeError("timestep: a+da does not converge onto the checkpoint scale factor %r (a=%r, da=%r)" % (a_chk, self.a, da))
        self.a_mid = f32(self.a + f32(da / f32(2)))
        self.dt, self.da = f32(dt), da
        self.tau, self.t = f32(self.tau + dt), f32(self.t + dt)
        self.a = f32(self.a + da)
        self.istep += 1
        return self.dt_old, self.dt, self.a_mid

    def after_checkpoint(self):                                # cafcube.f90:40-42
        self.cur_checkpoint += 1
        self.checkpoint_step = False
        self.dt = f32(0)
