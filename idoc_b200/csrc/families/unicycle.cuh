// Example of a user-defined iLQR problem family (the reference's ilqr.Problem takes Python callables `dynamics`,
// `cost`, `costf` and differentiates them with JAX, idoc/ilqr.py:94-158; here a family is the same three functions written
// once over a generic scalar S, and the device forward-mode AD of csrc/ad.cuh supplies every derivative the solver and
// the implicit VJP need).  Drop a file like this one into csrc/families/, end it with IDOC_FAMILY(<name>, <struct>), rebuild
// (python -c "import __graft_entry__ as g; g.build()"), and `idoc_b200.ilqr.Problem(family="<name>", ...)` works: fused
// solve + fused implicit VJP in fp32 and fp64, optional control limits.
//
// Kinematic unicycle driven to a goal position: state x = (px, py, phi), controls u = (v, omega),
//   px' = px + h v cos(phi),  py' = py + h v sin(phi),  phi' = phi + h omega
//   l  = 1/2 (w_p |p - goal|^2 + w_phi phi^2 + w_v v^2 + w_om omega^2),   lf = 1/2 (f_p |p - goal|^2 + f_phi phi^2)
// theta = [h, goal_x, goal_y, w_p, w_phi, w_v, w_om, f_p, f_phi]
#pragma once

struct UnicycleF {
  static constexpr int N = 3, M = 2, TD = 9;
  template <class S>
  __device__ static void dyn(const S* th, const S* x, const S* u, S* f) {
    f[0] = x[0] + th[0] * u[0] * mcos(x[2]);
    f[1] = x[1] + th[0] * u[0] * msin(x[2]);
    f[2] = x[2] + th[0] * u[1];
  }
  template <class S>
  __device__ static S cost(const S* th, const S* x, const S* u) {
    const S ex = x[0] - th[1], ey = x[1] - th[2];
    return S(0.5) * (th[3] * (ex * ex + ey * ey) + th[4] * x[2] * x[2] + th[5] * u[0] * u[0] + th[6] * u[1] * u[1]);
  }
  template <class S>
  __device__ static S costf(const S* th, const S* x) {
    const S ex = x[0] - th[1], ey = x[1] - th[2];
    return S(0.5) * (th[7] * (ex * ex + ey * ey) + th[8] * x[2] * x[2]);
  }
};
IDOC_FAMILY(unicycle, UnicycleF)
