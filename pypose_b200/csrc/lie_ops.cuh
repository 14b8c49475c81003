// lie_ops.cuh — per-op functors of the LieTensor op family: one group element in registers in,
// one out.  Host+device (see lie_math.cuh for why); the CUDA shells live in lie_kernels.cuh.
#pragma once
#include "lie_math.cuh"

namespace b200pose {

// ----------------------------------------------------------------------------
// op functors
// ----------------------------------------------------------------------------
#define OP_HEADER(NIN_, NOUT_, DI0_, DI1_, DI2_, DO0_, DO1_)                                    \
  using T = T_;                                                                                 \
  static constexpr int NIN = NIN_, NOUT = NOUT_, DI0 = DI0_, DI1 = DI1_, DI2 = DI2_, DO0 = DO0_, \
                       DO1 = DO1_;

// x(K) -> X(D)
template <class G, typename T_> struct OpExpFwd {
  OP_HEADER(1, 1, G::K, 1, 1, G::D, 1)
  static LM_HD void apply(const T* x, const T*, const T*, T* X, T*) {
    store_elem<G, T>(X, g_exp<G, T>(load_tang<G, T>(x)));
  }
};
// x(K), gX(D) -> gx(K) = gX[:K] @ Jl(x)
template <class G, typename T_> struct OpExpBwd {
  OP_HEADER(2, 1, G::K, G::D, 1, G::K, 1)
  static LM_HD void apply(const T* x, const T* g, const T*, T* gx, T*) {
    store_tang<G, T>(gx, jl_t_apply<G, T>(load_tang<G, T>(x), load_tang<G, T>(g)));
  }
};
// X(D) -> x(K)
template <class G, typename T_> struct OpLogFwd {
  OP_HEADER(1, 1, G::D, 1, 1, G::K, 1)
  static LM_HD void apply(const T* X, const T*, const T*, T* x, T*) {
    store_tang<G, T>(x, g_log<G, T>(load_elem<G, T>(X)));
  }
};
// out(K), g(K) -> gX(D) = [g @ Jl^-1(out), 0]
template <class G, typename T_> struct OpLogBwd {
  OP_HEADER(2, 1, G::K, G::K, 1, G::D, 1)
  static LM_HD void apply(const T* x, const T* g, const T*, T* gX, T*) {
    store_tang_pad<G, T>(gX, jlinv_t_apply<G, T>(load_tang<G, T>(x), load_tang<G, T>(g)));
  }
};
template <class G, typename T_> struct OpInvFwd {
  OP_HEADER(1, 1, G::D, 1, 1, G::D, 1)
  static LM_HD void apply(const T* X, const T*, const T*, T* Y, T*) {
    store_elem<G, T>(Y, g_inv<G, T>(load_elem<G, T>(X)));
  }
};
// Y(D), gY(D) -> gX(D) = [-gY[:K] @ Adj(Y), 0]   (op.py:944-949)
template <class G, typename T_> struct OpInvBwd {
  OP_HEADER(2, 1, G::D, G::D, 1, G::D, 1)
  static LM_HD void apply(const T* Y, const T* g, const T*, T* gX, T*) {
    store_tang_pad<G, T>(gX, t_neg(g_adj_t<G, T>(load_elem<G, T>(Y), load_tang<G, T>(g))));
  }
};
template <class G, typename T_> struct OpMulFwd {
  OP_HEADER(2, 1, G::D, G::D, 1, G::D, 1)
  static LM_HD void apply(const T* X, const T* Y, const T*, T* Z, T*) {
    store_elem<G, T>(Z, g_mul<G, T>(load_elem<G, T>(X), load_elem<G, T>(Y)));
  }
};
// X(D), gZ(D) -> gX(D) = [gZ[:K], 0], gY(D) = [gZ[:K] @ Adj(X), 0]   (op.py:845-852)
template <class G, typename T_> struct OpMulBwd {
  OP_HEADER(2, 2, G::D, G::D, 1, G::D, G::D)
  static LM_HD void apply(const T* X, const T* g, const T*, T* gX, T* gY) {
    Tang<T> gt = load_tang<G, T>(g);
    store_tang_pad<G, T>(gX, gt);
    store_tang_pad<G, T>(gY, g_adj_t<G, T>(load_elem<G, T>(X), gt));
  }
};
template <class G, typename T_> struct OpActFwd {
  OP_HEADER(2, 1, G::D, 3, 1, 3, 1)
  static LM_HD void apply(const T* X, const T* p, const T*, T* o, T*) {
    st3(o, g_act<G, T>(load_elem<G, T>(X), ld3(p)));
  }
};
// X(D), out(3), g(3) -> gX(D) = [g @ ActJac(out), 0], gp(3) = g @ M[:3,:3]   (op.py:534-542 etc.)
template <class G, typename T_> struct OpActBwd {
  OP_HEADER(3, 2, G::D, 3, 3, G::D, 3)
  static LM_HD void apply(const T* X, const T* out, const T* g, T* gX, T* gp) {
    Elem<T> e = load_elem<G, T>(X);
    V3<T> o = ld3(out), gg = ld3(g);
    Tang<T> t;
    t.tau = gg; t.phi = cross(o, gg); t.sigma = dot(gg, o);
    store_tang_pad<G, T>(gX, t);
    V3<T> r = qrot_t(e.q, gg);
    if (has_s<G>::v) r = e.s * r;
    st3(gp, r);
  }
};
template <class G, typename T_> struct OpAct4Fwd {
  OP_HEADER(2, 1, G::D, 4, 1, 4, 1)
  static LM_HD void apply(const T* X, const T* p, const T*, T* o, T*) {
    Elem<T> e = load_elem<G, T>(X);
    V3<T> r = qrot(e.q, ld3(p));
    if (has_s<G>::v) r = e.s * r;
    if (has_t<G>::v) r = r + p[3] * e.t;
    st3(o, r);
    o[3] = p[3];
  }
};
// X(D), out(4), g(4) -> gX(D) = [g @ Act4Jac(out), 0], gp(4) = g @ Matrix4x4(X)   (op.py:638-722)
template <class G, typename T_> struct OpAct4Bwd {
  OP_HEADER(3, 2, G::D, 4, 4, G::D, 4)
  static LM_HD void apply(const T* X, const T* out, const T* g, T* gX, T* gp) {
    Elem<T> e = load_elem<G, T>(X);
    V3<T> o = ld3(out), gg = ld3(g);
    Tang<T> t;
    t.tau = out[3] * gg; t.phi = cross(o, gg); t.sigma = dot(gg, o);
    store_tang_pad<G, T>(gX, t);
    V3<T> r = qrot_t(e.q, gg);
    if (has_s<G>::v) r = e.s * r;
    st3(gp, r);
    gp[3] = has_t<G>::v ? dot(gg, e.t) + g[3] : g[3];
  }
};
template <class G, typename T_> struct OpAdjFwd {
  OP_HEADER(2, 1, G::D, G::K, 1, G::K, 1)
  static LM_HD void apply(const T* X, const T* a, const T*, T* o, T*) {
    store_tang<G, T>(o, g_adj<G, T>(load_elem<G, T>(X), load_tang<G, T>(a)));
  }
};
// X(D), out(K), g(K) -> gX(D) = [-g @ ad(out), 0], ga(K) = g @ Adj(X)   (op.py:742-748 etc.)
template <class G, typename T_> struct OpAdjBwd {
  OP_HEADER(3, 2, G::D, G::K, G::K, G::D, G::K)
  static LM_HD void apply(const T* X, const T* out, const T* g, T* gX, T* ga) {
    Tang<T> gt = load_tang<G, T>(g);
    store_tang_pad<G, T>(gX, t_neg(ad_t_apply<G, T>(load_tang<G, T>(out), gt)));
    store_tang<G, T>(ga, g_adj_t<G, T>(load_elem<G, T>(X), gt));
  }
};
// Adj(X^-1) a   (op.py:1024-1030 etc.)
template <class G, typename T_> struct OpAdjTFwd {
  OP_HEADER(2, 1, G::D, G::K, 1, G::K, 1)
  static LM_HD void apply(const T* X, const T* a, const T*, T* o, T*) {
    store_tang<G, T>(o, g_adj<G, T>(g_inv<G, T>(load_elem<G, T>(X)), load_tang<G, T>(a)));
  }
};
// X(D), a(K), g(K) -> ga(K) = Adj(X) g, gX(D) = [-a @ ad(ga), 0]   (op.py:1038-1044 etc.)
template <class G, typename T_> struct OpAdjTBwd {
  OP_HEADER(3, 2, G::D, G::K, G::K, G::D, G::K)
  static LM_HD void apply(const T* X, const T* a, const T* g, T* gX, T* ga) {
    Tang<T> gat = g_adj<G, T>(load_elem<G, T>(X), load_tang<G, T>(g));
    store_tang<G, T>(ga, gat);
    store_tang_pad<G, T>(gX, t_neg(ad_t_apply<G, T>(gat, load_tang<G, T>(a))));
  }
};
// Jl^-1(Log X) p
template <class G, typename T_> struct OpJinvpFwd {
  OP_HEADER(2, 1, G::D, G::K, 1, G::K, 1)
  static LM_HD void apply(const T* X, const T* p, const T*, T* o, T*) {
    Tang<T> x = g_log<G, T>(load_elem<G, T>(X));
    store_tang<G, T>(o, jlinv_apply_g<G, T>(x, load_tang<G, T>(p)));
  }
};
// so3 right Jacobian (lietensor.py:343-351): Jr = I - c1 K + c2 K^2, identity when theta <= eps
template <typename T_> struct OpSo3Jr {
  OP_HEADER(1, 1, 3, 1, 1, 9, 1)
  static LM_HD void apply(const T* x, const T*, const T*, T* J, T*) {
    V3<T> phi = ld3(x);
    RotCoef<T> r = rot_coef(phi);
    T c1 = r.c1, c2 = r.c2;
    if (!(r.theta > num<T>::eps)) { c1 = T(0); c2 = T(0); }
    T xx = phi.x * phi.x, yy = phi.y * phi.y, zz = phi.z * phi.z;
    T xy = phi.x * phi.y, xz = phi.x * phi.z, yz = phi.y * phi.z;
    // K^2 = phi phi^T - theta^2 I
    J[0] = T(1) - c2 * (yy + zz);        J[1] = c1 * phi.z + c2 * xy;          J[2] = -c1 * phi.y + c2 * xz;
    J[3] = -c1 * phi.z + c2 * xy;        J[4] = T(1) - c2 * (xx + zz);         J[5] = c1 * phi.x + c2 * yz;
    J[6] = c1 * phi.y + c2 * xz;         J[7] = -c1 * phi.x + c2 * yz;         J[8] = T(1) - c2 * (xx + yy);
  }
};


// X-macro over the 17 ops every group implements: X(op_name, OpTemplate, NIN, NOUT, uses_algebra_prefix)
#define B200_FOR_EACH_GROUP_OP(X) \
  X(exp_fwd, OpExpFwd, 1, 1, 1)   \
  X(exp_bwd, OpExpBwd, 2, 1, 1)   \
  X(log_fwd, OpLogFwd, 1, 1, 0)   \
  X(log_bwd, OpLogBwd, 2, 1, 0)   \
  X(inv_fwd, OpInvFwd, 1, 1, 0)   \
  X(inv_bwd, OpInvBwd, 2, 1, 0)   \
  X(mul_fwd, OpMulFwd, 2, 1, 0)   \
  X(mul_bwd, OpMulBwd, 2, 2, 0)   \
  X(act_fwd, OpActFwd, 2, 1, 0)   \
  X(act_bwd, OpActBwd, 3, 2, 0)   \
  X(act4_fwd, OpAct4Fwd, 2, 1, 0) \
  X(act4_bwd, OpAct4Bwd, 3, 2, 0) \
  X(adj_fwd, OpAdjFwd, 2, 1, 0)   \
  X(adj_bwd, OpAdjBwd, 3, 2, 0)   \
  X(adjt_fwd, OpAdjTFwd, 2, 1, 0) \
  X(adjt_bwd, OpAdjTBwd, 3, 2, 0) \
  X(jinvp_fwd, OpJinvpFwd, 2, 1, 0)

}  // namespace b200pose
