// tpe_kernels.cuh -- v4 hot path: thread-per-env x chain-per-warp.
//
// Round-1 profiles showed the warp-per-env kernels spend ~56 k WARP instructions per env-substep on ~45 k THREAD
// instructions of essential work (4.7 of 32 lanes active, the tree is 9 levels deep and <= 5 bodies wide).  v4 turns the
// mapping around:
//   * one CTA = 32 envs; lane l of every warp owns env (32*blockIdx.x + l)  -> every instruction serves 32 envs;
//   * the 4 warps of the CTA are the 4 slots of the host list schedule of the kinematic tree (chain_host.hpp): at step t
//     warp w processes body sched[t][w] for its 32 envs; articulated inertias travel in registers along a chain and through a
//     shared-memory stash at junctions; a CTA barrier separates the steps (critical path 9 body-steps instead of 24);
//   * the SM's shared memory holds the hot per-env arrays of exactly these 32 envs in [word][lane] layout (bank = lane: conflict
//     free by construction); colder arrays (contacts, line-search rows, velocities) live in an L2-resident global scratch in
//     [word][env] layout (coalesced 128-B lines);
//   * body indices are warp-uniform, so the model table sits in __constant__ memory and is read through the uniform path.
// Mathematics identical to physics.cuh / warp_kernels.cuh (ABA, row-space line search).
#pragma once
#include "chain_model.cuh"
#include "dev_model.cuh"

#define TPE_MAXT 16
#define TPE_WARPS 4
#define TPE_SOLVER_MAXITER 12
#define TPE_LS_MAXITER 24

struct TpeTable {
  ChainConsts K;
  ChainEntry e[TPE_MAXT][TPE_WARPS];
};
__constant__ TpeTable c_tpe;

// ---- compile-time layouts (words per env)
template <int NB_, int NV_, int NE_>
struct TCfg {
  static constexpr int NB = NB_, NV = NV_, NQ = NV_ + 1, NU = NV_ - 6, NE = NE_;
  // shared memory
  static constexpr int qpos = 0, qvel = qpos + NQ, tau = qvel + NV, qacc = tau + NU, xpos = qacc + NV, quat = xpos + 3 * NB, ax = quat + 4 * NB,
                       U = ax + 3 * NV, Dinv = U + 6 * NV, u = Dinv + NV, acc = u + NV, stash = acc + 6 * NB,
                       red = stash + 27 * NE, smem_words = red + 4 * TPE_WARPS + 8;
  // global scratch
  static constexpr int g_act = 0, g_qstar = g_act + NU, g_spdab = g_qstar + NV, g_vel = g_spdab + NV, g_ab = g_vel + 6 * NB, g_acc2 = g_ab + 6 * NB,
                       g_ct1 = g_acc2 + 6 * NB, g_cpos = g_ct1 + 3 * NB, g_cD = g_cpos + 12 * NB, g_caref = g_cD + 4 * NB, g_cphi = g_caref + 16 * NB,
                       g_cr = g_cphi + 16 * NB, g_cdl = g_cr + 16 * NB, g_cflag = g_cdl + 16 * NB, g_lD = g_cflag + 4 * NB, g_laref = g_lD + NV,
                       g_lphi = g_laref + NV, g_lr = g_lphi + NV, g_ldl = g_lr + NV, g_lflag = g_ldl + NV, g_pb = g_lflag + NV, g_words = g_pb + 6 * NB;
};

struct TpeCtx {
  float* sm;           // CTA shared memory, [word][32]
  float* gs;           // global scratch, [word][npad]
  size_t npad;
  int lane, warp, genv;
  bool live;
};
#define TSM(x, off, i) ((x).sm[((off) + (i)) * 32 + (x).lane])
#define TGS(x, off, i) ((x).gs[(size_t)((off) + (i)) * (x).npad + (x).genv])
#define TGSI(x, off, i) (((int*)(x).gs)[(size_t)((off) + (i)) * (x).npad + (x).genv])

__device__ __forceinline__ V3 t_ld3s(const TpeCtx& x, int off, int i) { return v3(TSM(x, off, i), TSM(x, off, i + 1), TSM(x, off, i + 2)); }
__device__ __forceinline__ void t_st3s(const TpeCtx& x, int off, int i, V3 v) { TSM(x, off, i) = v.x; TSM(x, off, i + 1) = v.y; TSM(x, off, i + 2) = v.z; }
__device__ __forceinline__ S6 t_ld6s(const TpeCtx& x, int off, int i) { return s6(t_ld3s(x, off, i), t_ld3s(x, off, i + 3)); }
__device__ __forceinline__ void t_st6s(const TpeCtx& x, int off, int i, S6 v) { t_st3s(x, off, i, v.a); t_st3s(x, off, i + 3, v.l); }
__device__ __forceinline__ V3 t_ld3g(const TpeCtx& x, int off, int i) { return v3(TGS(x, off, i), TGS(x, off, i + 1), TGS(x, off, i + 2)); }
__device__ __forceinline__ void t_st3g(const TpeCtx& x, int off, int i, V3 v) { TGS(x, off, i) = v.x; TGS(x, off, i + 1) = v.y; TGS(x, off, i + 2) = v.z; }
__device__ __forceinline__ S6 t_ld6g(const TpeCtx& x, int off, int i) { return s6(t_ld3g(x, off, i), t_ld3g(x, off, i + 3)); }
__device__ __forceinline__ void t_st6g(const TpeCtx& x, int off, int i, S6 v) { t_st3g(x, off, i, v.a); t_st3g(x, off, i + 3, v.l); }

__device__ __forceinline__ void t_sincos(float x, float* s, float* c) {
  float k = rintf(x * 0.63661977236758134f);
  float r = fmaf(k, -1.5703125f, x);
  r = fmaf(k, -4.837512969970703125e-4f, r);
  r = fmaf(k, -7.549789954891882e-8f, r);
  float z = r * r;
  float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * r, r);
  float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z, fmaf(-0.5f, z, 1.0f));
  int q = ((int)k) & 3;
  float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ float t_impedance(float pm) {
  const ChainConsts& K = c_tpe.K;
  float x = fabsf(pm) / fmaxf(K.solimp[2], 1e-15f);
  if (x >= 1.f) return K.solimp[1];
  if (x <= 0.f) return K.solimp[0];
  float y, pw = K.solimp[4];
  if (pw == 2.0f) y = (x <= K.solimp[3]) ? K.imp_a * x * x : 1.f - K.imp_b * (1.f - x) * (1.f - x);
  else if (pw < 1.0000001f && pw > 0.9999999f) y = x;
  else y = (x <= K.solimp[3]) ? K.imp_a * __powf(x, pw) : 1.f - K.imp_b * __powf(1.f - x, pw);
  return K.solimp[0] + y * (K.solimp[1] - K.solimp[0]);
}

__device__ __forceinline__ S6 t_wrench(V3 cp, V3 t1, int k) {
  const ChainConsts& K = c_tpe.K;
  V3 n = ld3(K.plane_n);
  V3 t = (k < 2) ? t1 : cross(n, t1);
  float sg = (k & 1) ? -K.mu : K.mu;
  V3 dir = n + sg * t;
  return s6(cross(cp, dir), dir);
}

// motion subspace of dof k of entry e (about the root origin)
template <class C>
__device__ __forceinline__ S6 t_dofS(const TpeCtx& x, const ChainEntry& e, int k) {
  V3 a = t_ld3s(x, C::ax, 3 * (e.dofadr + k));
  if (e.kind == CH_KIND_ROOT6 && k < 3) return s6(v3(0.f, 0.f, 0.f), a);
  return s6(a, cross(t_ld3s(x, C::xpos, 3 * e.body), a));
}

template <class C>
__device__ __forceinline__ void t_rigid10(const TpeCtx& x, const ChainEntry& e, float* r10) {
  int b = e.body;
  Q4 q; q.w = TSM(x, C::quat, 4 * b); q.x = TSM(x, C::quat, 4 * b + 1); q.y = TSM(x, C::quat, 4 * b + 2); q.z = TSM(x, C::quat, 4 * b + 3);
  float R[9];
  q2mat(q, R);
  const float* in = e.inertia;
  float m = e.mass;
  V3 r = t_ld3s(x, C::xpos, 3 * b) + mrot(R, ld3(e.ipos));
  float Il[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, Tm[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Tm[3 * i + j] = R[3 * i] * Il[j] + R[3 * i + 1] * Il[3 + j] + R[3 * i + 2] * Il[6 + j];
  float rr = dot(r, r);
  r10[0] = m; r10[1] = m * r.x; r10[2] = m * r.y; r10[3] = m * r.z;
  r10[4] = Tm[0] * R[0] + Tm[1] * R[1] + Tm[2] * R[2] + m * (rr - r.x * r.x);
  r10[5] = Tm[3] * R[3] + Tm[4] * R[4] + Tm[5] * R[5] + m * (rr - r.y * r.y);
  r10[6] = Tm[6] * R[6] + Tm[7] * R[7] + Tm[8] * R[8] + m * (rr - r.z * r.z);
  r10[7] = Tm[0] * R[3] + Tm[1] * R[4] + Tm[2] * R[5] - m * r.x * r.y;
  r10[8] = Tm[0] * R[6] + Tm[1] * R[7] + Tm[2] * R[8] - m * r.x * r.z;
  r10[9] = Tm[3] * R[6] + Tm[4] * R[7] + Tm[5] * R[8] - m * r.y * r.z;
}

// per-env sum over the 4 warps (all threads of the CTA call this)
template <class C>
__device__ __forceinline__ float t_blocksum(const TpeCtx& x, float v) {
  x.sm[(C::red + x.warp) * 32 + x.lane] = v;
  __syncthreads();
  float s = x.sm[(C::red + 0) * 32 + x.lane] + x.sm[(C::red + 1) * 32 + x.lane] + x.sm[(C::red + 2) * 32 + x.lane] + x.sm[(C::red + 3) * 32 + x.lane];
  __syncthreads();
  return s;
}

// ------------------------------------------------------------------ kinematics / velocities / bias forces
template <class C>
__device__ __noinline__ void t_fk(const TpeCtx& x, bool vel) {
  const ChainConsts& K = c_tpe.K;
  for (int t = K.T - 1; t >= 0; t--) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb >= 0 && x.live) {
      int b = e.body;
      Q4 qc; V3 xp; S6 v, ab;
      v = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)); ab = v;
      if (e.kind == CH_KIND_ROOT6) {
        qc.w = TSM(x, C::qpos, 3); qc.x = TSM(x, C::qpos, 4); qc.y = TSM(x, C::qpos, 5); qc.z = TSM(x, C::qpos, 6);
        qc = qnormalize(qc);
        TSM(x, C::qpos, 3) = qc.w; TSM(x, C::qpos, 4) = qc.x; TSM(x, C::qpos, 5) = qc.y; TSM(x, C::qpos, 6) = qc.z;
        xp = v3(0.f, 0.f, 0.f);
        float R[9];
        q2mat(qc, R);
        t_st3s(x, C::ax, 0, v3(1.f, 0.f, 0.f)); t_st3s(x, C::ax, 3, v3(0.f, 1.f, 0.f)); t_st3s(x, C::ax, 6, v3(0.f, 0.f, 1.f));
        V3 c0 = v3(R[0], R[3], R[6]), c1 = v3(R[1], R[4], R[7]), c2 = v3(R[2], R[5], R[8]);
        t_st3s(x, C::ax, 9, c0); t_st3s(x, C::ax, 12, c1); t_st3s(x, C::ax, 15, c2);
        if (vel) {
          V3 vl = t_ld3s(x, C::qvel, 0), w = TSM(x, C::qvel, 3) * c0 + TSM(x, C::qvel, 4) * c1 + TSM(x, C::qvel, 5) * c2;
          v = s6(w, vl);
          ab = s6(v3(0.f, 0.f, 0.f), v3(-K.grav[0], -K.grav[1], -K.grav[2]) + cross(vl, w));
        }
      } else {
        int p = e.par_body;
        Q4 qp; qp.w = TSM(x, C::quat, 4 * p); qp.x = TSM(x, C::quat, 4 * p + 1); qp.y = TSM(x, C::quat, 4 * p + 2); qp.z = TSM(x, C::quat, 4 * p + 3);
        float Rp[9];
        q2mat(qp, Rp);
        xp = t_ld3s(x, C::xpos, 3 * p) + mrot(Rp, ld3(e.bpos));
        Q4 qb; qb.w = e.bquat[0]; qb.x = e.bquat[1]; qb.y = e.bquat[2]; qb.z = e.bquat[3];
        qc = qmul(qp, qb);
        if (vel) { v = t_ld6g(x, C::g_vel, 6 * p); ab = t_ld6g(x, C::g_ab, 6 * p); }
        for (int k = 0; k < e.ndof; k++) {
          int d = e.dofadr + k;
          V3 al = ld3(e.axis + 3 * k);
          V3 a = qrot(qc, al);
          t_st3s(x, C::ax, 3 * d, a);
          if (vel) {
            S6 S = s6(a, cross(xp, a));
            float qd = TSM(x, C::qvel, d);
            ab = ab + qd * cross_motion(v, S);
            v = v + qd * S;
          }
          float sn, cs;
          t_sincos(0.5f * TSM(x, C::qpos, d + 1), &sn, &cs);
          Q4 qj; qj.w = cs; qj.x = al.x * sn; qj.y = al.y * sn; qj.z = al.z * sn;
          qc = qmul(qc, qj);
        }
        qc = qnormalize(qc);
      }
      t_st3s(x, C::xpos, 3 * b, xp);
      TSM(x, C::quat, 4 * b) = qc.w; TSM(x, C::quat, 4 * b + 1) = qc.x; TSM(x, C::quat, 4 * b + 2) = qc.y; TSM(x, C::quat, 4 * b + 3) = qc.z;
      if (vel) {
        t_st6g(x, C::g_vel, 6 * b, v); t_st6g(x, C::g_ab, 6 * b, ab);
        float r10[10];
        t_rigid10<C>(x, e, r10);
        t_st6g(x, C::g_pb, 6 * b, rb_mul(r10, ab) + cross_force(v, rb_mul(r10, v)));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ ABA inward sweep (flags / tmode / dmode as in chain_kernels.cuh)
#define T_INERTIA 1
#define T_FORCE 2
#define T_PB 4
#define T_CONTACTS 8
template <class C>
__device__ __noinline__ void t_inward(const TpeCtx& x, bool run, int flags, int tmode, int dmode) {
  const ChainConsts& K = c_tpe.K;
  float A[21];
  S6 p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
#pragma unroll
  for (int j = 0; j < 21; j++) A[j] = 0.f;
  const bool inertia = flags & T_INERTIA, force = flags & T_FORCE;
  const unsigned rm = __ballot_sync(0xffffffffu, run);   // lanes that execute the body-ops below (uniform e.pb)
  for (int t = 0; t < K.T; t++) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb >= 0 && run) {
      int b = e.body;
      if (!e.carry_in) {
#pragma unroll
        for (int j = 0; j < 21; j++) A[j] = 0.f;
        p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
      }
      if (inertia) {
        float r10[10], B[21];
        t_rigid10<C>(x, e, r10);
        rb_expand(r10, B);
#pragma unroll
        for (int j = 0; j < 21; j++) A[j] += B[j];
      }
      if (force && (flags & T_PB)) p = p + t_ld6g(x, C::g_pb, 6 * b);
      if ((flags & T_CONTACTS) && e.geom >= 0) {
        int any = 0;
#pragma unroll
        for (int s = 0; s < 4; s++) any |= TGSI(x, C::g_cflag, 4 * b + s);
        if (__any_sync(rm, (any & 30) != 0)) {
          V3 t1 = t_ld3g(x, C::g_ct1, 3 * b);
          for (int s = 0; s < 4; s++) {
            int fl = TGSI(x, C::g_cflag, 4 * b + s);
            if (!(fl & 1)) fl = 0;
            if (!__any_sync(rm, (fl & 30) != 0)) continue;
            V3 cp = t_ld3g(x, C::g_cpos, 12 * b + 3 * s);
            float D = TGS(x, C::g_cD, 4 * b + s);
            for (int k = 0; k < 4; k++) {
              bool on = (fl & (2 << k)) != 0;
              if (!__any_sync(rm, on)) continue;
              S6 xw = t_wrench(cp, t1, k);
              float xv[6] = {xw.a.x, xw.a.y, xw.a.z, xw.l.x, xw.l.y, xw.l.z};
              float Dk = on ? D : 0.f;
              if (!on) {
#pragma unroll
                for (int j = 0; j < 6; j++) xv[j] = 0.f;   // unused slots hold stale data: select, never multiply
              }
              if (inertia) sym_rank1(A, xv, -Dk);
              if (force && on) p = p - (D * TGS(x, C::g_caref, 16 * b + 4 * s + k)) * xw;
            }
          }
        }
      }
      for (int j = 0; j < 3; j++) {
        int ed = e.in_edge[j];
        if (ed < 0) continue;
        int so = C::stash + 27 * (ed - CH_EDGE_MBOX);
        if (inertia) {
#pragma unroll
          for (int i = 0; i < 21; i++) A[i] += TSM(x, so, i);
        }
        if (force) p = p + t_ld6s(x, so, 21);
      }
      for (int k = e.ndof - 1; k >= 0; k--) {
        int d = e.dofadr + k;
        S6 S = t_dofS<C>(x, e, k);
        float s[6] = {S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z}, Uv[6], di;
        const bool hinge = e.kind == CH_KIND_HINGE;
        int lf = (hinge && (tmode == 0 || dmode == 0)) ? TGSI(x, C::g_lflag, d) : 0;
        if (inertia) {
          sym_mul(A, s, Uv);
          float D = hinge ? e.arm[k] : 0.f;
          if (dmode == 1) D += hinge ? K.h * e.kd[k] : 0.f;
          else if (lf & 4) D += TGS(x, C::g_lD, d);
#pragma unroll
          for (int j = 0; j < 6; j++) D = fmaf(s[j], Uv[j], D);
          di = 1.0f / D;
#pragma unroll
          for (int j = 0; j < 6; j++) TSM(x, C::U, 6 * d + j) = Uv[j];
          TSM(x, C::Dinv, d) = di;
          sym_rank1(A, Uv, di);
        } else {
#pragma unroll
          for (int j = 0; j < 6; j++) Uv[j] = TSM(x, C::U, 6 * d + j);
          di = TSM(x, C::Dinv, d);
        }
        if (force) {
          float tin = 0.f;
          if (hinge) {
            if (tmode == 0) {
              tin = TSM(x, C::tau, d - 6);
              if (lf & 4) tin += ((lf & 1) ? 1.f : -1.f) * TGS(x, C::g_lD, d) * TGS(x, C::g_laref, d);
            } else if (tmode == 2) {
              float tgt = fmaf(TGS(x, C::g_act, d - 6), e.ascale[k], e.aoffset[k]);
              float err = TSM(x, C::qpos, d + 1) + TSM(x, C::qvel, d) * K.h - tgt;
              tin = -e.kp[k] * err - e.kd[k] * TSM(x, C::qvel, d);
            }
          }
          float uu = tin - dot6(S, p);
          TSM(x, C::u, d) = uu;
          p = p + (uu * di) * s6(v3(Uv[0], Uv[1], Uv[2]), v3(Uv[3], Uv[4], Uv[5]));
        }
      }
      if (!e.carry_out && e.out_edge >= 0) {
        int so = C::stash + 27 * (e.out_edge - CH_EDGE_MBOX);
        if (inertia) {
#pragma unroll
          for (int i = 0; i < 21; i++) TSM(x, so, i) = A[i];
        }
        if (force) t_st6s(x, so, 21, p);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ ABA outward sweep.  which: 0 qacc | 1 qstar | 2 spdab ; mode 1: accumulate S*qacc only
template <class C>
__device__ __noinline__ void t_outward(const TpeCtx& x, bool run, int which, int mode) {
  const ChainConsts& K = c_tpe.K;
  for (int t = K.T - 1; t >= 0; t--) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb >= 0 && run) {
      int b = e.body;
      S6 a = (e.kind == CH_KIND_ROOT6) ? s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)) : t_ld6s(x, C::acc, 6 * e.par_body);
      for (int k = 0; k < e.ndof; k++) {
        int d = e.dofadr + k;
        S6 S = t_dofS<C>(x, e, k);
        float qdd;
        if (mode == 1) qdd = TSM(x, C::qacc, d);
        else {
          qdd = TSM(x, C::Dinv, d) * (TSM(x, C::u, d) - dot6(t_ld6s(x, C::U, 6 * d), a));
          if (which == 0) TSM(x, C::qacc, d) = qdd; else if (which == 1) TGS(x, C::g_qstar, d) = qdd; else TGS(x, C::g_spdab, d) = qdd;
        }
        a = a + qdd * S;
      }
      t_st6s(x, C::acc, 6 * b, a);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ collision + limit rows for this warp's bodies; returns mask bits and row count
template <class C>
__device__ __noinline__ unsigned long long t_collide(const TpeCtx& x, int* nrows_out) {
  const ChainConsts& K = c_tpe.K;
  unsigned long long mask = 0ull;
  int nrows = 0;
  if (x.live) {
    V3 n = ld3(K.plane_n);
    float root_h = dot(n, t_ld3s(x, C::qpos, 0) - ld3(K.plane_pos));
    for (int t = 0; t < K.T; t++) {
      const ChainEntry& e = c_tpe.e[t][x.warp];
      if (e.pb < 0) continue;
      int b = e.body;
      if (e.kind == CH_KIND_HINGE) {
        for (int k = 0; k < e.ndof; k++) {
          int d = e.dofadr + k, f = 0;
          if ((e.limited >> k) & 1) {
            float q = TSM(x, C::qpos, d + 1), dlo = q - e.range[2 * k], dhi = e.range[2 * k + 1] - q, dist = 0.f, sg = 0.f;
            if (dlo < 0.f) { dist = dlo; sg = 1.f; f = 1; }
            else if (dhi < 0.f) { dist = dhi; sg = -1.f; f = 2; }
            if (f) {
              float imp = t_impedance(dist);
              TGS(x, C::g_lD, d) = 1.0f / fmaxf((1.f - imp) / imp * e.diw0[k], 1e-15f);
              TGS(x, C::g_laref, d) = -K.B * sg * TSM(x, C::qvel, d) - K.K * imp * dist;
              nrows++;
            }
          }
          TGSI(x, C::g_lflag, d) = f;
        }
      }
#pragma unroll
      for (int s = 0; s < 4; s++) TGSI(x, C::g_cflag, 4 * b + s) = 0;
      if (e.geom < 0) continue;
      Q4 q; q.w = TSM(x, C::quat, 4 * b); q.x = TSM(x, C::quat, 4 * b + 1); q.y = TSM(x, C::quat, 4 * b + 2); q.z = TSM(x, C::quat, 4 * b + 3);
      float R[9];
      q2mat(q, R);
      V3 c = t_ld3s(x, C::xpos, 3 * b) + mrot(R, ld3(e.gpos));
      float d0 = root_h + dot(n, c);
      int cnt = 0;
      V3 t1 = ld3(K.t1_default);
      float dist_s[4];
      V3 cp_s[4];
      if (e.gtype == SMPLSIM_GEOM_CAPSULE || e.gtype == SMPLSIM_GEOM_SPHERE) {
        V3 axw = mrot(R, v3(e.gmat[2], e.gmat[5], e.gmat[8]));
        float rad = e.gsize[0], hl = (e.gtype == SMPLSIM_GEOM_CAPSULE) ? e.gsize[1] : 0.f, na = dot(n, axw);
        int nend = (e.gtype == SMPLSIM_GEOM_CAPSULE) ? 2 : 1;
        if (nend == 2) {
          t1 = axw - na * n;
          float nn = sqrtf(dot(t1, t1));
          t1 = (nn < 1e-15f) ? v3(1.f, 0.f, 0.f) : (1.0f / nn) * t1;
        }
        for (int en = 0; en < nend; en++) {
          float sg = en ? -hl : hl, dist = d0 + sg * na - rad;
          if (dist > K.margin) continue;
          cp_s[cnt] = c + sg * axw - (rad + 0.5f * dist) * n; dist_s[cnt] = dist; cnt++;
        }
      } else {
        for (int i = 0; i < 8 && cnt < 4; i++) {
          V3 vl = v3((i & 1) ? e.gsize[0] : -e.gsize[0], (i & 2) ? e.gsize[1] : -e.gsize[1], (i & 4) ? e.gsize[2] : -e.gsize[2]);
          V3 w = mrot(R, mrot(e.gmat, vl));
          float l = dot(n, w);
          if (d0 + l > K.margin || l > 0.f) continue;
          float dist = d0 + l;
          cp_s[cnt] = c + w - (0.5f * dist) * n; dist_s[cnt] = dist; cnt++;
        }
      }
      if (cnt) {
        t_st3g(x, C::g_ct1, 3 * b, t1);
        S6 v = t_ld6g(x, C::g_vel, 6 * b);
        for (int s = 0; s < cnt; s++) {
          t_st3g(x, C::g_cpos, 12 * b + 3 * s, cp_s[s]);
          float pm = dist_s[s] - K.margin, imp = t_impedance(pm);
          float R0 = fmaxf((1.f - imp) / imp * (e.tran_iw0 + K.mu * K.mu * e.tran_iw0), 1e-15f);
          float R1 = R0 / fmaxf(K.impratio, 1e-15f), mu = K.mu * sqrtf(R1 / R0);
          TGS(x, C::g_cD, 4 * b + s) = 1.0f / (2.f * mu * mu * R0);
          float kterm = K.K * imp * pm;
          for (int k = 0; k < 4; k++) TGS(x, C::g_caref, 16 * b + 4 * s + k) = -K.B * dot6(t_wrench(cp_s[s], t1, k), v) - kterm;
          TGSI(x, C::g_cflag, 4 * b + s) = 1;
        }
        mask |= 1ull << (e.geom + 1);
        nrows += 4 * cnt;
      }
    }
  }
  *nrows_out = nrows;
  return mask;
}

// ------------------------------------------------------------------ constraint rows (this warp's bodies)
// op 0: working set := (r<0) at acc / qacc (warm start)
// op 1: rs at acc / qstar -> cdl := rs ; returns per-thread "same as working set"
// op 2: adopt  (r := rs, phi := -D rs on used set, set := rs<0)
// op 3: prepare line search (cdl := rs - r ; g1, g2 accumulated into out)
// op 4: line-search sums at step al -> out[2], out[3]
// op 5: apply step al (r += al d, phi += al (phis - phi), set := r<0)
template <class C>
__device__ __noinline__ bool t_rows(const TpeCtx& x, bool run, int op, float al, float* out4) {
  const ChainConsts& K = c_tpe.K;
  bool same = true;
  float g1 = 0.f, g2 = 0.f, s1 = 0.f, s2 = 0.f;
  if (run) {
    for (int t = 0; t < K.T; t++) {
      const ChainEntry& e = c_tpe.e[t][x.warp];
      if (e.pb < 0) continue;
      int b = e.body;
      if (e.geom >= 0) {
        S6 a;
        V3 t1;
        bool loaded = false;
        for (int s = 0; s < 4; s++) {
          int fl = TGSI(x, C::g_cflag, 4 * b + s);
          if (!(fl & 1)) continue;
          if (!loaded && op <= 1) { a = t_ld6s(x, C::acc, 6 * b); t1 = t_ld3g(x, C::g_ct1, 3 * b); loaded = true; }
          float D = TGS(x, C::g_cD, 4 * b + s);
          int nf = 1;
          V3 cp = (op <= 1) ? t_ld3g(x, C::g_cpos, 12 * b + 3 * s) : v3(0.f, 0.f, 0.f);
          for (int k = 0; k < 4; k++) {
            int ri = 16 * b + 4 * s + k;
            if (op <= 1) {
              float r = dot6(t_wrench(cp, t1, k), a) - TGS(x, C::g_caref, ri);
              if (op == 1) TGS(x, C::g_cdl, ri) = r;
              if (r < 0.f) nf |= 2 << k;
            } else if (op == 2) {
              float rs = TGS(x, C::g_cdl, ri);
              TGS(x, C::g_cr, ri) = rs; TGS(x, C::g_cphi, ri) = (fl & (2 << k)) ? -D * rs : 0.f;
              if (rs < 0.f) nf |= 2 << k;
            } else if (op == 3) {
              float rs = TGS(x, C::g_cdl, ri), d = rs - TGS(x, C::g_cr, ri), ph = TGS(x, C::g_cphi, ri), phs = (fl & (2 << k)) ? -D * rs : 0.f;
              TGS(x, C::g_cdl, ri) = d;
              g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
            } else {
              float r = TGS(x, C::g_cr, ri), d = TGS(x, C::g_cdl, ri), v = fmaf(al, d, r);
              if (op == 4) { if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); } }
              else {
                float rs = r + d, ph = TGS(x, C::g_cphi, ri), phs = (fl & (2 << k)) ? -D * rs : 0.f;
                TGS(x, C::g_cr, ri) = v; TGS(x, C::g_cphi, ri) = fmaf(al, phs - ph, ph);
                if (v < 0.f) nf |= 2 << k;
              }
            }
          }
          if (op == 0 || op == 2 || op == 5) TGSI(x, C::g_cflag, 4 * b + s) = nf;
          else if (op == 1 && nf != fl) same = false;
        }
      }
      if (e.kind == CH_KIND_HINGE) {
        for (int k = 0; k < e.ndof; k++) {
          int d = e.dofadr + k, f = TGSI(x, C::g_lflag, d);
          if (!(f & 3)) continue;
          float sg = (f & 1) ? 1.f : -1.f, D = TGS(x, C::g_lD, d);
          int nf = f & 3;
          if (op <= 1) {
            float r = sg * (op == 0 ? TSM(x, C::qacc, d) : TGS(x, C::g_qstar, d)) - TGS(x, C::g_laref, d);
            if (op == 1) TGS(x, C::g_ldl, d) = r;
            if (r < 0.f) nf |= 4;
            if (op == 0) TGSI(x, C::g_lflag, d) = nf; else if (nf != f) same = false;
          } else if (op == 2) {
            float rs = TGS(x, C::g_ldl, d);
            TGS(x, C::g_lr, d) = rs; TGS(x, C::g_lphi, d) = (f & 4) ? -D * rs : 0.f;
            TGSI(x, C::g_lflag, d) = nf | ((rs < 0.f) ? 4 : 0);
          } else if (op == 3) {
            float rs = TGS(x, C::g_ldl, d), dd = rs - TGS(x, C::g_lr, d), ph = TGS(x, C::g_lphi, d), phs = (f & 4) ? -D * rs : 0.f;
            TGS(x, C::g_ldl, d) = dd;
            g1 = fmaf(dd, ph, g1); g2 = fmaf(dd, phs - ph, g2);
          } else {
            float r = TGS(x, C::g_lr, d), dd = TGS(x, C::g_ldl, d), v = fmaf(al, dd, r);
            if (op == 4) { if (v < 0.f) { s1 = fmaf(D * v, dd, s1); s2 = fmaf(D * dd, dd, s2); } }
            else {
              float rs = r + dd, ph = TGS(x, C::g_lphi, d), phs = (f & 4) ? -D * rs : 0.f;
              TGS(x, C::g_lr, d) = v; TGS(x, C::g_lphi, d) = fmaf(al, phs - ph, ph);
              TGSI(x, C::g_lflag, d) = nf | ((v < 0.f) ? 4 : 0);
            }
          }
        }
      }
    }
  }
  if (out4) { out4[0] = g1; out4[1] = g2; out4[2] = s1; out4[3] = s2; }
  return same;
}

// this warp's share of an axpy over its dofs / bodies:  qacc += al (qstar - qacc) ; acc2 += al (acc - acc2)   (al = 1: copy)
template <class C>
__device__ __noinline__ void t_advance(const TpeCtx& x, bool doit, float al) {
  const ChainConsts& K = c_tpe.K;
  if (!doit) return;
  for (int t = 0; t < K.T; t++) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb < 0) continue;
    for (int k = 0; k < e.ndof; k++) {
      int d = e.dofadr + k;
      float q = TSM(x, C::qacc, d), qs = TGS(x, C::g_qstar, d);
      TSM(x, C::qacc, d) = (al == 1.f) ? qs : fmaf(al, qs - q, q);
    }
    for (int j = 0; j < 6; j++) {
      float a2 = TGS(x, C::g_acc2, 6 * e.body + j);
      TGS(x, C::g_acc2, 6 * e.body + j) = (al == 1.f) ? TSM(x, C::acc, 6 * e.body + j) : fmaf(al, TSM(x, C::acc, 6 * e.body + j) - a2, a2);
    }
  }
}

// ------------------------------------------------------------------ constraint solve (CTA-uniform control flow; per-env predicates)
template <class C>
__device__ __noinline__ int t_solve(const TpeCtx& x, bool any_rows) {
  bool plain = x.live && !any_rows;
  if (__syncthreads_or(plain)) {
    t_inward<C>(x, plain, T_INERTIA | T_FORCE | T_PB, 0, 0);
    t_outward<C>(x, plain, 0, 0);
  }
  bool run = x.live && any_rows;
  if (!__syncthreads_or(run)) return 0;
  t_outward<C>(x, run, 0, 1);
  t_rows<C>(x, run, 0, 0.f, nullptr);
  bool have_point = false;
  int it = 0, iters = 0;
  float o4[4];
  for (; it < TPE_SOLVER_MAXITER; it++) {
    if (!__syncthreads_or(run)) break;
    t_inward<C>(x, run, T_INERTIA | T_FORCE | T_PB | T_CONTACTS, 0, 0);
    t_outward<C>(x, run, 1, 0);
    bool mysame = t_rows<C>(x, run, 1, 0.f, nullptr);
    bool same = t_blocksum<C>(x, mysame ? 0.f : 1.f) == 0.f;
    bool fin = run && same, adopt = run && !same && !have_point, lsrch = run && !same && have_point;
    if (adopt) t_rows<C>(x, true, 2, 0.f, nullptr);
    t_advance<C>(x, fin || adopt, 1.f);
    if (fin) { run = false; iters = it; }
    if (adopt) have_point = true;
    if (__syncthreads_or(lsrch)) {
      t_rows<C>(x, lsrch, 3, 0.f, o4);
      float g1 = t_blocksum<C>(x, o4[0]), g2 = t_blocksum<C>(x, o4[1]);
      t_rows<C>(x, lsrch, 4, 0.f, o4);
      float s1 = t_blocksum<C>(x, o4[2]), s2;
      float f0 = g1 + s1, al = 0.f, lo = 0.f, hi = -1.f, tol = 1e-6f * fabsf(f0);
      bool searching = lsrch && (f0 < -W_LS_NOISE * (fabsf(g1) + fabsf(s1)));
      if (searching) al = 1.f;
      for (int ls = 0; ls < TPE_LS_MAXITER; ls++) {
        if (!__syncthreads_or(searching)) break;
        t_rows<C>(x, searching, 4, al, o4);
        s1 = t_blocksum<C>(x, o4[2]); s2 = t_blocksum<C>(x, o4[3]);
        if (searching) {
          float f = g1 + al * g2 + s1, fp = g2 + s2;
          if (fabsf(f) <= tol) searching = false;
          else {
            if (f < 0.f) lo = al; else hi = al;
            float an = (fp > 0.f) ? al - f / fp : -1.f;
            if (!(an > lo) || (hi > 0.f && !(an < hi))) an = (hi > 0.f) ? 0.5f * (lo + hi) : 2.f * al;
            an = fminf(an, W_LS_MAXSTEP);
            if (an == al) searching = false; else al = an;
          }
        }
      }
      bool step = lsrch && (al > 0.f);
      if (lsrch && !step) { run = false; iters = it; }
      t_rows<C>(x, step, 5, al, nullptr);
      t_advance<C>(x, step, al);
    }
    __syncthreads();
  }
  if (run) iters = it;
  return iters;
}

template <class C>
__device__ __noinline__ void t_spd_prepare(const TpeCtx& x) {
  t_inward<C>(x, x.live, T_INERTIA | T_FORCE | T_PB, 1, 1);
  t_outward<C>(x, x.live, 2, 0);
}

template <class C>
__device__ __noinline__ void t_torque(const TpeCtx& x) {
  const ChainConsts& K = c_tpe.K;
  int mode = K.cfg.control_mode;
  if (mode == SMPLSIM_CTRL_UHC_PD) {
    t_inward<C>(x, x.live, T_FORCE, 2, 1);
    t_outward<C>(x, x.live, 1, 0);
  }
  if (x.live) {
    for (int t = 0; t < K.T; t++) {
      const ChainEntry& e = c_tpe.e[t][x.warp];
      if (e.pb < 0 || e.kind != CH_KIND_HINGE) continue;
      for (int k = 0; k < e.ndof; k++) {
        int d = e.dofadr + k;
        float a = TGS(x, C::g_act, d - 6), tq, q = TSM(x, C::qpos, d + 1), qd = TSM(x, C::qvel, d);
        if (mode == SMPLSIM_CTRL_TORQUE) tq = a * e.ascale[k];
        else {
          float tgt = fmaf(a, e.ascale[k], e.aoffset[k]);
          if (mode == SMPLSIM_CTRL_PD) tq = -e.kp[k] * (q - tgt) - e.kd[k] * qd;
          else tq = -e.kp[k] * (q + qd * K.h - tgt) - e.kd[k] * (qd + (TGS(x, C::g_spdab, d) + TGS(x, C::g_qstar, d)) * K.h);
        }
        TSM(x, C::tau, d - 6) = fminf(fmaxf(tq, -e.tlim[k]), e.tlim[k]);
      }
    }
  }
  __syncthreads();
}

// semi-implicit Euler over this warp's dofs; the warp owning the root returns the root displacement
template <class C>
__device__ __noinline__ V3 t_integrate(const TpeCtx& x) {
  const ChainConsts& K = c_tpe.K;
  V3 disp = v3(0.f, 0.f, 0.f);
  if (x.live) {
    float h = K.h;
    for (int t = 0; t < K.T; t++) {
      const ChainEntry& e = c_tpe.e[t][x.warp];
      if (e.pb < 0) continue;
      for (int k = 0; k < e.ndof; k++) {
        int d = e.dofadr + k;
        float v = fmaf(h, TSM(x, C::qacc, d), TSM(x, C::qvel, d));
        TSM(x, C::qvel, d) = v;
        if (e.kind == CH_KIND_HINGE) TSM(x, C::qpos, d + 1) = fmaf(h, v, TSM(x, C::qpos, d + 1));
      }
      if (e.kind == CH_KIND_ROOT6) {
        disp = h * t_ld3s(x, C::qvel, 0);
        t_st3s(x, C::qpos, 0, t_ld3s(x, C::qpos, 0) + disp);
        V3 w = t_ld3s(x, C::qvel, 3);
        float n = sqrtf(dot(w, w)), ang = n * h;
        Q4 q; q.w = TSM(x, C::qpos, 3); q.x = TSM(x, C::qpos, 4); q.y = TSM(x, C::qpos, 5); q.z = TSM(x, C::qpos, 6);
        if (ang > 0.f) {
          float sn, cs;
          t_sincos(0.5f * ang, &sn, &cs);
          float s = sn / n;
          Q4 dq; dq.w = cs; dq.x = w.x * s; dq.y = w.y * s; dq.z = w.z * s;
          q = qmul(q, dq);
        }
        q = qnormalize(q);
        TSM(x, C::qpos, 3) = q.w; TSM(x, C::qpos, 4) = q.x; TSM(x, C::qpos, 5) = q.y; TSM(x, C::qpos, 6) = q.z;
      }
    }
  }
  __syncthreads();
  return disp;
}

// ====================================================================================================================
// env-level kernels (v4): one CTA of 4 warps = 32 envs
// ====================================================================================================================
struct TpeStepArgs {
  SmplsimState st;
  SmplsimAux aux;
  const float* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  float* gs;
  size_t npad;
  int n, nsub, mode;
};
struct TpeResetArgs {
  SmplsimState st;
  SmplsimAux aux;
  const uint8_t* mask;
  const float* qpos0;
  const float* qvel0;
  float* obs;
  float* gs;
  size_t npad;
  int n, init_mode;
};

struct TpeTask { float target[4]; int change, cur_t, recov; uint32_t rng; };   // per-env task scalars (registers of warp 0)

// what: bit0 qpos/qvel ; bit1 warm start -> qacc ; bit2 action -> g_act ; bit3 action -> tau
template <class C>
__device__ __noinline__ void t_load(const TpeCtx& x, const float* qp, const float* qv, const float* qw, const float* a, int what) {
  if (x.live) {
    if (what & 1) {
      for (int i = x.warp; i < C::NQ; i += TPE_WARPS) TSM(x, C::qpos, i) = qp[i];
      for (int i = x.warp; i < C::NV; i += TPE_WARPS) TSM(x, C::qvel, i) = qv[i];
    }
    if (what & 2) for (int i = x.warp; i < C::NV; i += TPE_WARPS) TSM(x, C::qacc, i) = qw[i];
    if (what & 4) for (int i = x.warp; i < C::NU; i += TPE_WARPS) TGS(x, C::g_act, i) = a[i];
    if (what & 8) for (int i = x.warp; i < C::NU; i += TPE_WARPS) TSM(x, C::tau, i) = a[i];
  }
  __syncthreads();
}
template <class C>
__device__ __noinline__ void t_store(const TpeCtx& x, float* qp, float* qv, float* qw) {
  __syncthreads();
  if (x.live) {
    for (int i = x.warp; i < C::NQ; i += TPE_WARPS) qp[i] = TSM(x, C::qpos, i);
    for (int i = x.warp; i < C::NV; i += TPE_WARPS) qv[i] = TSM(x, C::qvel, i);
    if (qw) for (int i = x.warp; i < C::NV; i += TPE_WARPS) qw[i] = TSM(x, C::qacc, i);
  }
  __syncthreads();   // every warp reads every warp's dofs here: nobody may start integrating before all are done
}

__device__ __forceinline__ void t_reset_task(TpeTask& k, int env) {
  const SmplsimEnvCfg& c = c_tpe.K.cfg;
  if (c.task == SMPLSIM_TASK_NONE) return;
  uint32_t r[4];
  philox4x32(k.rng, (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
  k.rng += 1;
  if (c.task == SMPLSIM_TASK_SPEED) k.target[0] = (float)(c.tar_speed_max - c.tar_speed_min) * u01(r[0]) + (float)c.tar_speed_min;
  else if (c.task == SMPLSIM_TASK_REACH) {
    k.target[0] = (float)c.tar_dist_max * (2.0f * u01(r[0]) - 1.0f);
    k.target[1] = (float)c.tar_dist_max * (2.0f * u01(r[1]) - 1.0f);
    k.target[2] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[2]) + (float)c.tar_height_min;
  } else k.target[0] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[0]) + (float)c.tar_height_min;
  k.change = k.cur_t + rand_range(r[3], c.change_steps_min, c.change_steps_max);
}
__device__ __forceinline__ void t_task_io(const TpeCtx& x, TpeTask& k, const SmplsimState& st, bool store) {
  if (!x.live || x.warp != 0) return;
  int env = x.genv;
  if (!store) {
    for (int j = 0; j < 4; j++) k.target[j] = st.task_target[4 * env + j];
    k.change = st.task_change_step[env]; k.cur_t = st.progress[env]; k.recov = st.recovery[env]; k.rng = st.rng_counter[env];
  } else {
    for (int j = 0; j < 4; j++) st.task_target[4 * env + j] = k.target[j];
    st.task_change_step[env] = k.change; st.progress[env] = k.cur_t; st.recovery[env] = k.recov; st.rng_counter[env] = k.rng;
  }
}

struct TpeFwd { unsigned long long mask; int iters; };

template <class C>
__device__ __noinline__ V3 t_substeps(const TpeCtx& x, int nsub, int raw, TpeFwd* fo, const SmplsimState& st, bool write_fwd, bool prep_last) {
  const ChainConsts& K = c_tpe.K;
  const bool spd = (K.cfg.control_mode == SMPLSIM_CTRL_UHC_PD), stale = K.cfg.spd_stale != 0;
  V3 disp = v3(0.f, 0.f, 0.f);
  for (int s = 0; s < nsub; s++) {
    bool did_fk = false;
    if (!raw) {
      if (spd && !stale) { t_fk<C>(x, true); t_spd_prepare<C>(x); did_fk = true; }
      t_torque<C>(x);
    }
    if (!did_fk) t_fk<C>(x, true);
    int nrows = 0;
    unsigned long long m = t_collide<C>(x, &nrows);
    // per-env OR of the contact masks / row counts over the 4 warps
    unsigned* ru = (unsigned*)x.sm;
    ru[(C::red + x.warp) * 32 + x.lane] = (unsigned)(m & 0xffffffffull);
    ru[(C::red + 4 + x.warp) * 32 + x.lane] = (unsigned)(m >> 32);
    ru[(C::red + 8 + x.warp) * 32 + x.lane] = (unsigned)nrows;
    __syncthreads();
    unsigned lo = 0, hi = 0, nr = 0;
    for (int w = 0; w < TPE_WARPS; w++) { lo |= ru[(C::red + w) * 32 + x.lane]; hi |= ru[(C::red + 4 + w) * 32 + x.lane]; nr += ru[(C::red + 8 + w) * 32 + x.lane]; }
    __syncthreads();
    fo->mask = ((unsigned long long)hi << 32) | lo;
    fo->iters = t_solve<C>(x, nr > 0);
    if (s == nsub - 1) {
      if (x.live) {
        for (int t = 0; t < K.T; t++) {       // framelinvel / frameangvel of the last forward pass (quirk Q2), parked in g_acc2
          const ChainEntry& e = c_tpe.e[t][x.warp];
          if (e.pb < 0) continue;
          S6 v = t_ld6g(x, C::g_vel, 6 * e.body);
          t_st3g(x, C::g_acc2, 6 * e.body, v.l + cross(v.a, t_ld3s(x, C::xpos, 3 * e.body)));
          t_st3g(x, C::g_acc2, 6 * e.body + 3, v.a);
        }
      }
      if (write_fwd) t_store<C>(x, st.qpos_fwd + (size_t)x.genv * C::NQ, st.qvel_fwd + (size_t)x.genv * C::NV, nullptr);
    }
    if (spd && stale && !raw && (s < nsub - 1 || prep_last)) t_spd_prepare<C>(x);
    disp = disp + t_integrate<C>(x);
  }
  return disp;
}

__device__ __forceinline__ Q4 t_heading_inv(Q4 root) {
  if (!c_tpe.K.cfg.upright_start) { Q4 bc; bc.w = 0.5f; bc.x = -0.5f; bc.y = -0.5f; bc.z = -0.5f; root = qmul(root, bc); }
  V3 rd = qrot_ref(root, v3(1.f, 0.f, 0.f));
  float hd = atan2f(rd.y, rd.x), sn, cs;
  t_sincos(-0.5f * hd, &sn, &cs);
  Q4 h; h.w = cs; h.x = 0.f; h.y = 0.f; h.z = sn;
  return qnormalize(h);
}

// compute_observations: every warp writes the entries of its own bodies; warp 0 adds the root / task entries
template <class C>
__device__ __noinline__ void t_write_obs(const TpeCtx& x, const TpeTask& k, float* ob) {
  const ChainConsts& K = c_tpe.K;
  if (!x.live || !ob) return;
  Q4 r0; r0.w = TSM(x, C::quat, 0); r0.x = TSM(x, C::quat, 1); r0.y = TSM(x, C::quat, 2); r0.z = TSM(x, C::quat, 3);
  Q4 hq = t_heading_inv(r0);
  int nb = K.nb, o = K.cfg.root_height_obs ? 1 : 0, o_rot = o + 3 * (nb - 1), o_vel = o_rot + 6 * nb;
  for (int t = 0; t < K.T; t++) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb < 0) continue;
    int b = e.body;
    if (b > 0) st3(ob + o + 3 * (b - 1), qrot_ref(hq, t_ld3s(x, C::xpos, 3 * b)));
    Q4 q; q.w = TSM(x, C::quat, 4 * b); q.x = TSM(x, C::quat, 4 * b + 1); q.y = TSM(x, C::quat, 4 * b + 2); q.z = TSM(x, C::quat, 4 * b + 3);
    Q4 lq = qmul(hq, q);
    st3(ob + o_rot + 6 * b, qrot_ref(lq, v3(1.f, 0.f, 0.f)));
    st3(ob + o_rot + 6 * b + 3, qrot_ref(lq, v3(0.f, 0.f, 1.f)));
    if (K.cfg.self_obs_v == 2) {
      st3(ob + o_vel + 3 * b, qrot_ref(hq, t_ld3g(x, C::g_acc2, 6 * b)));
      st3(ob + o_vel + 3 * nb + 3 * b, qrot_ref(hq, t_ld3g(x, C::g_acc2, 6 * b + 3)));
    } else if (e.kind == CH_KIND_HINGE) {
      for (int kk = 0; kk < e.ndof; kk++) ob[o_vel + e.dofadr + kk] = TSM(x, C::qvel, e.dofadr + kk);
    }
  }
  if (x.warp == 0) {
    if (o) ob[0] = TSM(x, C::qpos, 2);
    if (K.cfg.self_obs_v == 1) {
      st3(ob + o_vel, qrot_ref(hq, t_ld3s(x, C::qvel, 0)));
      st3(ob + o_vel + 3, qrot_ref(hq, t_ld3s(x, C::qvel, 3)));
    }
    int ot = K.self_obs_dim;
    Q4 rq; rq.w = TSM(x, C::qpos, 3); rq.x = TSM(x, C::qpos, 4); rq.y = TSM(x, C::qpos, 5); rq.z = TSM(x, C::qpos, 6);
    if (K.cfg.task == SMPLSIM_TASK_SPEED) {
      V3 d = qrot_ref(t_heading_inv(rq), v3(1.f, 0.f, 0.f));
      ob[ot] = d.x; ob[ot + 1] = d.y; ob[ot + 2] = k.target[0];
    } else if (K.cfg.task == SMPLSIM_TASK_REACH) st3(ob + ot, qrot_ref(t_heading_inv(rq), ld3(k.target) - t_ld3s(x, C::qpos, 0)));
    else if (K.cfg.task == SMPLSIM_TASK_GETUP) ob[ot] = k.target[0];
  }
}

template <class C>
__device__ __noinline__ void t_write_aux(const TpeCtx& x, const SmplsimAux& aux, const TpeFwd& fo) {
  const ChainConsts& K = c_tpe.K;
  if (!x.live) return;
  int env = x.genv;
  V3 root = t_ld3s(x, C::qpos, 0);
  for (int t = 0; t < K.T; t++) {
    const ChainEntry& e = c_tpe.e[t][x.warp];
    if (e.pb < 0) continue;
    int b = e.body;
    size_t bi = (size_t)env * K.nb + b;
    if (aux.xpos) st3(aux.xpos + bi * 3, t_ld3s(x, C::xpos, 3 * b) + root);
    if (aux.xquat) for (int j = 0; j < 4; j++) aux.xquat[bi * 4 + j] = TSM(x, C::quat, 4 * b + j);
    if (aux.body_linvel) st3(aux.body_linvel + bi * 3, t_ld3g(x, C::g_acc2, 6 * b));
    if (aux.body_angvel) st3(aux.body_angvel + bi * 3, t_ld3g(x, C::g_acc2, 6 * b + 3));
    for (int k = 0; k < e.ndof; k++) {
      int d = e.dofadr + k;
      if (aux.qacc) aux.qacc[(size_t)env * K.nv + d] = TSM(x, C::qacc, d);
      if (aux.ctrl && e.kind == CH_KIND_HINGE) aux.ctrl[(size_t)env * K.nu + d - 6] = TSM(x, C::tau, d - 6);
    }
  }
  if (x.warp == 0) {
    if (aux.contact_mask) aux.contact_mask[env] = fo.mask;
    if (aux.solver_iter) aux.solver_iter[env] = fo.iters;
  }
}

extern __shared__ float t_smem[];

template <class C>
__global__ void __launch_bounds__(128, 1) k_step4(TpeStepArgs a) {
  const ChainConsts& K = c_tpe.K;
  TpeCtx x;
  x.sm = t_smem; x.gs = a.gs; x.npad = a.npad; x.lane = threadIdx.x & 31; x.warp = threadIdx.x >> 5;
  x.genv = blockIdx.x * 32 + x.lane;
  x.live = x.genv < a.n;
  if (!x.live) x.genv = a.n - 1;          // safe addresses for predicated-off lanes
  size_t eo = (size_t)x.genv;
  const bool spd = (K.cfg.control_mode == SMPLSIM_CTRL_UHC_PD);
  if (spd && K.cfg.spd_stale && a.mode == 0) {
    t_load<C>(x, a.st.qpos_fwd + eo * C::NQ, a.st.qvel_fwd + eo * C::NV, nullptr, nullptr, 1);
    t_fk<C>(x, true);
    t_spd_prepare<C>(x);
  }
  t_load<C>(x, a.st.qpos + eo * C::NQ, a.st.qvel + eo * C::NV, a.st.qacc_warm + eo * C::NV, a.action + eo * C::NU, 1 | 2 | (a.mode == 0 ? 4 : 8));
  TpeTask tk;
  t_task_io(x, tk, a.st, false);
  if (a.mode == 0 && x.live && x.warp == 0 && K.cfg.task != SMPLSIM_TASK_NONE && tk.cur_t >= tk.change) t_reset_task(tk, x.genv);
  TpeFwd fo; fo.mask = 0ull; fo.iters = 0;
  V3 disp = t_substeps<C>(x, a.nsub, a.mode, &fo, a.st, true, false);
  // root displacement lives in the warp that owns the root: publish per env
  {
    float* rd = x.sm + (C::red) * 32;
    bool owns_root = false;
    for (int t = 0; t < K.T; t++) if (c_tpe.e[t][x.warp].pb >= 0 && c_tpe.e[t][x.warp].kind == CH_KIND_ROOT6) owns_root = true;
    if (owns_root) { rd[x.lane] = disp.x; rd[32 + x.lane] = disp.y; }
    __syncthreads();
    disp.x = rd[x.lane]; disp.y = rd[32 + x.lane];
    __syncthreads();
  }
  t_fk<C>(x, false);
  if (a.mode == 0) {
    if (x.warp == 0) tk.cur_t += 1;
    t_write_obs<C>(x, tk, a.obs ? a.obs + eo * K.obs_dim : nullptr);
    if (x.live && x.warp == 0) {
      const SmplsimEnvCfg& c = K.cfg;
      float rew = 0.f;
      if (c.task == SMPLSIM_TASK_SPEED) {
        float inv_dt = 1.0f / (K.h * (float)a.nsub), vx = disp.x * inv_dt, vy = disp.y * inv_dt, er = tk.target[0] - vx;
        rew = expf(-0.25f * (er * er + 0.1f * vy * vy));
      } else if (c.task == SMPLSIM_TASK_REACH) {
        V3 dl = ld3(tk.target) - (t_ld3s(x, C::xpos, 3 * c.reach_body) + t_ld3s(x, C::qpos, 0));
        rew = expf(-4.0f * dot(dl, dl));
      } else if (c.task == SMPLSIM_TASK_GETUP) { float er = tk.target[0] - TSM(x, C::qpos, 2); rew = expf(-4.0f * er * er); }
      int term = 0, trunc = 0, pass_time = tk.cur_t > c.episode_length;
      if (c.task == SMPLSIM_TASK_NONE) trunc = pass_time;
      else if (c.task == SMPLSIM_TASK_GETUP && tk.recov > 0) tk.recov -= 1;
      else { trunc = pass_time; term = (fo.mask & ~K.legal_mask) != 0ull; }
      if (a.reward) a.reward[x.genv] = rew;
      if (a.terminated) a.terminated[x.genv] = (uint8_t)term;
      if (a.truncated) a.truncated[x.genv] = (uint8_t)trunc;
    }
  }
  t_write_aux<C>(x, a.aux, fo);
  t_store<C>(x, a.st.qpos + eo * C::NQ, a.st.qvel + eo * C::NV, a.st.qacc_warm + eo * C::NV);
  if (a.mode == 0) t_task_io(x, tk, a.st, true);
}

template <class C>
__global__ void __launch_bounds__(128, 1) k_reset4(TpeResetArgs a) {
  const ChainConsts& K = c_tpe.K;
  const SmplsimEnvCfg& c = K.cfg;
  TpeCtx x;
  x.sm = t_smem; x.gs = a.gs; x.npad = a.npad; x.lane = threadIdx.x & 31; x.warp = threadIdx.x >> 5;
  x.genv = blockIdx.x * 32 + x.lane;
  x.live = x.genv < a.n;
  if (!x.live) x.genv = a.n - 1;
  if (x.live && a.mask && !a.mask[x.genv]) x.live = false;
  if (!__syncthreads_or(x.live)) return;
  size_t eo = (size_t)x.genv;
  int init = a.init_mode < 0 ? c.state_init : a.init_mode;
  TpeTask tk;
  t_task_io(x, tk, a.st, false);
  if (x.live && x.warp == 0) {
    if (c.task == SMPLSIM_TASK_GETUP) tk.recov = c.recovery_steps;
    if (!c.legacy_change_step) tk.cur_t = 0;
    t_reset_task(tk, x.genv);
  }
  if (x.live) {
    for (int i = x.warp; i < C::NQ; i += TPE_WARPS) TSM(x, C::qpos, i) = 0.f;
    for (int i = x.warp; i < C::NV; i += TPE_WARPS) { TSM(x, C::qvel, i) = 0.f; TSM(x, C::qacc, i) = 0.f; }
    for (int i = x.warp; i < C::NU; i += TPE_WARPS) { TSM(x, C::tau, i) = 0.f; TGS(x, C::g_act, i) = 0.f; }
  }
  __syncthreads();
  TpeFwd fo; fo.mask = 0ull; fo.iters = 0;
  if (init == SMPLSIM_INIT_MOCAP) t_load<C>(x, a.qpos0 + eo * C::NQ, a.qvel0 + eo * C::NV, nullptr, nullptr, 1);
  else {
    if (x.live && x.warp == 0) {
      if (init == SMPLSIM_INIT_DEFAULT) { TSM(x, C::qpos, 2) = 0.94f; for (int j = 3; j < 7; j++) TSM(x, C::qpos, j) = 0.5f; }
      else { TSM(x, C::qpos, 2) = 0.3f; TSM(x, C::qpos, 3) = 1.0f; }
    }
    __syncthreads();
  }
  if (init == SMPLSIM_INIT_FALL) {
    if (c.control_mode == SMPLSIM_CTRL_UHC_PD && c.spd_stale) { t_fk<C>(x, true); t_spd_prepare<C>(x); }
    int ngrp = (K.nu + 3) / 4;
    uint32_t base = 0;
    for (int k3 = 0; k3 < 3; k3++) {
      // every warp needs the env's rng counter: broadcast from warp 0 through shared memory
      unsigned* ru = (unsigned*)x.sm + (C::red) * 32;
      if (x.warp == 0) ru[x.lane] = tk.rng;
      __syncthreads();
      base = ru[x.lane];
      __syncthreads();
      if (x.live) {
        for (int i = x.warp; i < K.nu; i += TPE_WARPS) {
          uint32_t r[4];
          philox4x32(base + (uint32_t)(i >> 2), (uint32_t)x.genv, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
          TGS(x, C::g_act, i) = u01(r[i & 3]) - 0.5f;
        }
      }
      if (x.warp == 0) tk.rng = base + (uint32_t)ngrp;
      __syncthreads();
      t_substeps<C>(x, c.nsubsteps, 0, &fo, a.st, false, true);
    }
  }
  t_fk<C>(x, true);
  {
    int nrows = 0;
    unsigned long long m = t_collide<C>(x, &nrows);
    unsigned* ru = (unsigned*)x.sm;
    ru[(C::red + x.warp) * 32 + x.lane] = (unsigned)(m & 0xffffffffull);
    ru[(C::red + 4 + x.warp) * 32 + x.lane] = (unsigned)(m >> 32);
    __syncthreads();
    unsigned lo = 0, hi = 0;
    for (int w = 0; w < TPE_WARPS; w++) { lo |= ru[(C::red + w) * 32 + x.lane]; hi |= ru[(C::red + 4 + w) * 32 + x.lane]; }
    __syncthreads();
    fo.mask = ((unsigned long long)hi << 32) | lo;
  }
  if (x.live) {
    for (int t = 0; t < K.T; t++) {
      const ChainEntry& e = c_tpe.e[t][x.warp];
      if (e.pb < 0) continue;
      S6 v = t_ld6g(x, C::g_vel, 6 * e.body);
      t_st3g(x, C::g_acc2, 6 * e.body, v.l + cross(v.a, t_ld3s(x, C::xpos, 3 * e.body)));
      t_st3g(x, C::g_acc2, 6 * e.body + 3, v.a);
    }
  }
  if (x.warp == 0) tk.cur_t = 0;
  __syncthreads();
  t_write_obs<C>(x, tk, a.obs ? a.obs + eo * K.obs_dim : nullptr);
  t_write_aux<C>(x, a.aux, fo);
  t_store<C>(x, a.st.qpos + eo * C::NQ, a.st.qvel + eo * C::NV, a.st.qacc_warm + eo * C::NV);
  t_store<C>(x, a.st.qpos_fwd + eo * C::NQ, a.st.qvel_fwd + eo * C::NV, nullptr);
  t_task_io(x, tk, a.st, true);
}
