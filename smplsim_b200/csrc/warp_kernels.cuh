// warp_kernels.cuh -- v3 hot path: level-synchronous ABA stepper, LPE lanes per env (LPE = 32 or 16).
//
// Same mathematics as physics.cuh (v1).  What changed, driven by the round-1 ncu profiles
// (profiles/r1_k_step_v1.md): v1 was instruction-fetch bound (310 KB of SASS, "no_instruction" the top
// stall) at 8 warps/SM (26.8 KB of shared memory per env) with the constant table read through global
// memory.  v3 keeps one env per LPE-lane group with everything in shared memory, but
//   * every sweep is ONE non-inlined function with runtime flags (hot code fits the instruction cache);
//   * the shared-memory layout is compile-time (template on model size) and aliased: one U/Dinv array
//     serves the stable-PD factors and the solver (the SPD factors of state s_k are produced after the
//     solve of substep k and consumed by the torque pass of substep k+1 before anything overwrites
//     them), rigid inertias are rebuilt on the fly, line-search rows are recomputed from two
//     acceleration fields instead of being stored  ->  ~14 KB per env (SMPL);
//   * the per-body / per-dof / per-geom constants are staged once per CTA into shared memory;
//   * compact sincos; powf only on the non-default solimp path.
#pragma once
#include "dev_model.cuh"

#define W_FULL 0xffffffffu
#define W_SOLVER_MAXITER 12
#define W_LS_MAXITER 24
#define W_LS_NOISE 1e-4f     // line search: directional derivative below this fraction of its two cancelling parts = converged
#define W_LS_MAXSTEP 16.f    // line search: never extrapolate further than this multiple of the Newton step
#define W_MAXLIM 8

// ------------------------------------------------------------------ compile-time sizes / shared-memory layout
template <int NB_, int NV_, int NG_, int NS_, int LPE_>
struct WCfg {
  static constexpr int NB = NB_, NV = NV_, NQ = NV_ + 1, NU = NV_ - 6, NG = NG_, NS = NS_, LPE = LPE_, EPW = 32 / LPE_;
  static constexpr int r4(int n) { return (n + 3) & ~3; }
  // per-env scratch (words)
  static constexpr int qpos = 0, qvel = qpos + r4(NQ), act = qvel + r4(NV), tau = act + r4(NU), qacc = tau + r4(NU), qstar = qacc + r4(NV),
                       spdab = qstar + r4(NV), xpos = spdab + r4(NV), xquat = xpos + r4(3 * NB), ax = xquat + r4(4 * NB), vel = ax + r4(3 * NV),
                       pb = vel + r4(6 * NB), irb = pb + r4(6 * NB), IA = irb + r4(10 * NB), pA = IA + r4(21 * NB), U = pA + r4(6 * NB), Dinv = U + r4(6 * NV),
                       u = Dinv + r4(NV), acc = u + r4(NV), acc2 = acc + r4(6 * NB), cpos = acc2 + r4(6 * NB), cD = cpos + r4(3 * NS),
                       caref = cD + r4(NS), cphi = caref + r4(4 * NS), cflag = cphi + r4(4 * NS), ct1 = cflag + r4(NS), lim = ct1 + r4(3 * NG),
                       tsk = lim + r4(8 * W_MAXLIM + 4), total_ = tsk + 12;
  static constexpr int total = total_ | 4;   // env stride not a multiple of 32 words
  // obs staging aliases the IA accumulators (dead outside the sweeps)
  static constexpr int obs = IA;
};

// constants staged per CTA (float32 image of the model, exact sizes)
template <class C>
struct WModel {
  int nb, nv, nu, ng, nlevel, nslot, sched_T, rowpar, warmset, dirtypath;
  float ls_tol;
  int sched[SM_MAXSCHED][4], sched_nd[SM_MAXSCHED], sched_nc[SM_MAXSCHED], sched_ns[SM_MAXSCHED];
  int parent[C::NB], dofadr[C::NB], dofnum[C::NB];
  int level_adr[SM_MAXL + 1], level_list[C::NB], child_adr[C::NB + 1], child_list[C::NB];
  int bgeom_adr[C::NB + 1], bgeom_list[C::NG], gtype[C::NG], gbody[C::NG], slot_adr[C::NG + 1], slot_geom[C::NS], limited[C::NV];
  float bpos[C::NB][3], bquat[C::NB][4], mass[C::NB], ipos[C::NB][3], inertia[C::NB][6], tran_iw0[C::NB];
  float axis[C::NV][3], arm[C::NV], diw0[C::NV], range[C::NV][2];
  float kp[C::NV], kd[C::NV], tlim[C::NV], ascale[C::NV], aoffset[C::NV];
  float gpos[C::NG][3], gmat[C::NG][9], gsize[C::NG][3];
  float plane_pos[3], plane_n[3], t1_default[3], margin, mu, impratio, solimp[5], imp_a, imp_b, K, B, h, grav[3];
  unsigned long long legal_mask;
  SmplsimEnvCfg cfg;
  int obs_dim, self_obs_dim;
};

template <class C>
__device__ void w_stage_model(const DevModel* __restrict__ G, WModel<C>& M) {
  int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    M.nb = G->nb; M.nv = G->nv; M.nu = G->nu; M.ng = G->ng; M.nlevel = G->nlevel; M.nslot = G->nslot;
    for (int i = 0; i <= SM_MAXL; i++) M.level_adr[i] = G->level_adr[i];
    for (int i = 0; i < 3; i++) { M.plane_pos[i] = G->plane_pos[i]; M.plane_n[i] = G->plane_n[i]; M.t1_default[i] = G->t1_default[i]; M.grav[i] = G->grav[i]; }
    M.margin = G->margin; M.mu = G->mu; M.impratio = G->impratio;
    for (int i = 0; i < 5; i++) M.solimp[i] = G->solimp[i];
    M.imp_a = G->imp_a; M.imp_b = G->imp_b; M.K = G->K; M.B = G->B; M.h = G->h; M.legal_mask = G->legal_mask; M.cfg = G->cfg;
    M.obs_dim = G->obs_dim; M.self_obs_dim = G->self_obs_dim;
    M.sched_T = G->sched_T; M.rowpar = G->rowpar; M.warmset = G->warmset; M.dirtypath = G->dirtypath; M.ls_tol = G->ls_tol;
    for (int i = 0; i < SM_MAXSCHED; i++) {
      for (int k = 0; k < 4; k++) M.sched[i][k] = G->sched[i][k];
      M.sched_nd[i] = G->sched_nd[i]; M.sched_nc[i] = G->sched_nc[i]; M.sched_ns[i] = G->sched_ns[i];
    }
  }
  int nb = G->nb, nv = G->nv, ng = G->ng, ns = G->nslot;
  for (int b = tid; b < nb; b += nt) {
    M.parent[b] = G->parent[b]; M.dofadr[b] = G->dofadr[b]; M.dofnum[b] = G->dofnum[b]; M.level_list[b] = G->level_list[b];
    M.child_list[b] = G->child_list[b]; M.mass[b] = G->mass[b]; M.tran_iw0[b] = G->tran_iw0[b];
    for (int k = 0; k < 3; k++) { M.bpos[b][k] = G->bpos[b][k]; M.ipos[b][k] = G->ipos[b][k]; }
    for (int k = 0; k < 4; k++) M.bquat[b][k] = G->bquat[b][k];
    for (int k = 0; k < 6; k++) M.inertia[b][k] = G->inertia[b][k];
  }
  for (int b = tid; b <= nb; b += nt) { M.child_adr[b] = G->child_adr[b]; M.bgeom_adr[b] = G->bgeom_adr[b]; }
  for (int d = tid; d < nv; d += nt) {
    for (int k = 0; k < 3; k++) M.axis[d][k] = G->axis[d][k];
    M.arm[d] = G->arm[d]; M.diw0[d] = G->diw0[d]; M.range[d][0] = G->range[d][0]; M.range[d][1] = G->range[d][1]; M.limited[d] = G->limited[d];
    M.kp[d] = G->kp[d]; M.kd[d] = G->kd[d]; M.tlim[d] = G->tlim[d]; M.ascale[d] = G->ascale[d]; M.aoffset[d] = G->aoffset[d];
  }
  for (int g = tid; g < ng; g += nt) {
    M.gtype[g] = G->gtype[g]; M.gbody[g] = G->gbody[g]; M.bgeom_list[g] = G->bgeom_list[g];
    for (int k = 0; k < 3; k++) { M.gpos[g][k] = G->gpos[g][k]; M.gsize[g][k] = G->gsize[g][k]; }
    for (int k = 0; k < 9; k++) M.gmat[g][k] = G->gmat[g][k];
  }
  for (int g = tid; g <= ng; g += nt) M.slot_adr[g] = G->slot_adr[g];
  for (int s = tid; s < ns; s += nt) M.slot_geom[s] = G->slot_geom[s];
}

// lane context
struct WLane {
  int li;        // lane within the env group
  int lane;      // lane within the warp
  unsigned gmask;  // warp mask of this env's lanes
  bool live;
};

template <class C>
__device__ __forceinline__ float w_gsum(float v) {
#pragma unroll
  for (int o = C::LPE / 2; o > 0; o >>= 1) v += __shfl_xor_sync(W_FULL, v, o);
  return v;
}
template <class C>
__device__ __forceinline__ unsigned w_gor(unsigned v) {
#pragma unroll
  for (int o = C::LPE / 2; o > 0; o >>= 1) v |= __shfl_xor_sync(W_FULL, v, o);
  return v;
}
__device__ __forceinline__ bool w_gall(bool p, const WLane& w) { return (__ballot_sync(W_FULL, p) & w.gmask) == w.gmask; }
__device__ __forceinline__ bool w_gany(bool p, const WLane& w) { return (__ballot_sync(W_FULL, p) & w.gmask) != 0u; }

__device__ __forceinline__ void w_sincos(float x, float* s, float* c) {
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

template <class C>
__device__ __noinline__ float w_impedance(const WModel<C>& M, float pm) {
  float x = fabsf(pm) / fmaxf(M.solimp[2], 1e-15f);
  if (x >= 1.f) return M.solimp[1];
  if (x <= 0.f) return M.solimp[0];
  float y, pw = M.solimp[4];
  if (pw == 2.0f) y = (x <= M.solimp[3]) ? M.imp_a * x * x : 1.f - M.imp_b * (1.f - x) * (1.f - x);
  else if (pw < 1.0000001f && pw > 0.9999999f) y = x;
  else y = (x <= M.solimp[3]) ? M.imp_a * __powf(x, pw) : 1.f - M.imp_b * __powf(1.f - x, pw);
  return M.solimp[0] + y * (M.solimp[1] - M.solimp[0]);
}

template <class C>
__device__ __forceinline__ S6 w_dofS(const WModel<C>& M, const float* sm, int b, int k) {
  V3 a = ld3(sm + C::ax + 3 * (M.dofadr[b] + k));
  if (b == 0 && k < 3) return s6(v3(0.f, 0.f, 0.f), a);
  return s6(a, cross(ld3(sm + C::xpos + 3 * b), a));
}

template <class C>
__device__ __forceinline__ void w_rigid10(const WModel<C>& M, const float* sm, int b, float* r10) {
  const float* qq = sm + C::xquat + 4 * b;
  Q4 q; q.w = qq[0]; q.x = qq[1]; q.y = qq[2]; q.z = qq[3];
  float R[9];
  q2mat(q, R);
  const float* in = M.inertia[b];
  float m = M.mass[b];
  V3 r = ld3(sm + C::xpos + 3 * b) + mrot(R, ld3(M.ipos[b]));
  float Il[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, T[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) T[3 * i + j] = R[3 * i] * Il[j] + R[3 * i + 1] * Il[3 + j] + R[3 * i + 2] * Il[6 + j];
  float rr = dot(r, r);
  r10[0] = m; r10[1] = m * r.x; r10[2] = m * r.y; r10[3] = m * r.z;
  r10[4] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + m * (rr - r.x * r.x);
  r10[5] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + m * (rr - r.y * r.y);
  r10[6] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + m * (rr - r.z * r.z);
  r10[7] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - m * r.x * r.y;
  r10[8] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - m * r.x * r.z;
  r10[9] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - m * r.y * r.z;
}

template <class C>
__device__ __forceinline__ S6 w_wrench(const WModel<C>& M, V3 cp, V3 t1, int k) {
  V3 n = ld3(M.plane_n);
  V3 t = (k < 2) ? t1 : cross(n, t1);
  float sg = (k & 1) ? -M.mu : M.mu;
  V3 dir = n + sg * t;
  return s6(cross(cp, dir), dir);
}

// limit rows: compact list in shared memory.  entry e: [dof, side(+1/-1), D, aref, phi, flags(bit0 working set)] ; count at lim[0]
#define WLIM(sm, e, f) ((sm)[C::lim + 4 + 8 * (e) + (f)])

// ------------------------------------------------------------------ kinematics (+ velocities, bias forces): level-synchronous outward sweep
template <class C>
__device__ __noinline__ void w_fk(const WModel<C>& M, float* sm, const WLane& w, bool vel) {
  for (int lev = 0; lev < M.nlevel; lev++) {
    if (w.live) {
      for (int i = M.level_adr[lev] + w.li; i < M.level_adr[lev + 1]; i += C::LPE) {
        int b = M.level_list[i];
        Q4 qc; V3 x; S6 v, ab;
        v = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)); ab = v;
        float* qpos = sm + C::qpos;
        const float* qvel = sm + C::qvel;
        float* ax = sm + C::ax;
        if (b == 0) {
          qc.w = qpos[3]; qc.x = qpos[4]; qc.y = qpos[5]; qc.z = qpos[6];
          qc = qnormalize(qc);
          qpos[3] = qc.w; qpos[4] = qc.x; qpos[5] = qc.y; qpos[6] = qc.z;
          x = v3(0.f, 0.f, 0.f);
          float R[9];
          q2mat(qc, R);
          st3(ax + 0, v3(1.f, 0.f, 0.f)); st3(ax + 3, v3(0.f, 1.f, 0.f)); st3(ax + 6, v3(0.f, 0.f, 1.f));
          V3 c0 = v3(R[0], R[3], R[6]), c1 = v3(R[1], R[4], R[7]), c2 = v3(R[2], R[5], R[8]);
          st3(ax + 9, c0); st3(ax + 12, c1); st3(ax + 15, c2);
          if (vel) {
            V3 vl = ld3(qvel), wv = qvel[3] * c0 + qvel[4] * c1 + qvel[5] * c2;
            v = s6(wv, vl);
            ab = s6(v3(0.f, 0.f, 0.f), v3(-M.grav[0], -M.grav[1], -M.grav[2]) + cross(vl, wv));
          }
        } else {
          int p = M.parent[b];
          const float* qq = sm + C::xquat + 4 * p;
          Q4 qp; qp.w = qq[0]; qp.x = qq[1]; qp.y = qq[2]; qp.z = qq[3];
          float Rp[9];
          q2mat(qp, Rp);
          x = ld3(sm + C::xpos + 3 * p) + mrot(Rp, ld3(M.bpos[b]));
          Q4 qb; qb.w = M.bquat[b][0]; qb.x = M.bquat[b][1]; qb.y = M.bquat[b][2]; qb.z = M.bquat[b][3];
          qc = qmul(qp, qb);
          if (vel) { v = ld6(sm + C::vel + 6 * p); ab = ld6(sm + C::acc + 6 * p); }   // bias acceleration parked in acc during FK
          int d0 = M.dofadr[b], nd = M.dofnum[b];
          for (int k = 0; k < nd; k++) {
            int d = d0 + k;
            V3 al = ld3(M.axis[d]);
            V3 a = qrot(qc, al);
            st3(ax + 3 * d, a);
            if (vel) {
              S6 S = s6(a, cross(x, a));
              float qd = qvel[d];
              ab = ab + qd * cross_motion(v, S);
              v = v + qd * S;
            }
            float sn, cs;
            w_sincos(0.5f * qpos[d + 1], &sn, &cs);
            Q4 qj; qj.w = cs; qj.x = al.x * sn; qj.y = al.y * sn; qj.z = al.z * sn;
            qc = qmul(qc, qj);
          }
          qc = qnormalize(qc);
        }
        float* xq = sm + C::xquat + 4 * b;
        xq[0] = qc.w; xq[1] = qc.x; xq[2] = qc.y; xq[3] = qc.z;
        st3(sm + C::xpos + 3 * b, x);
        if (vel) {
          st6(sm + C::vel + 6 * b, v);
          st6(sm + C::acc + 6 * b, ab);
          float r10[10];
          w_rigid10(M, sm, b, r10);
#pragma unroll
          for (int j = 0; j < 10; j++) sm[C::irb + 10 * b + j] = r10[j];
          st6(sm + C::pb + 6 * b, rb_mul(r10, ab) + cross_force(v, rb_mul(r10, v)));
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ row-parallel sweeps (LPE == 32): 8 lanes per body, 4 bodies per step
// Lane r < 6 of an 8-lane group owns row r of the body's 6x6 articulated inertia (and component r of the bias force /
// acceleration); the 6-vector products become 8-lane shuffle reductions.  Bodies follow the host list schedule
// M.sched[t][slot] (children strictly before parents, <= 4 bodies per step) instead of the depth levels, so a sweep costs
// ~T x (a third of the per-body instructions) instead of nlevel x (the slowest body of the level).
__device__ __forceinline__ float w_sel6(int r, float a0, float a1, float a2, float a3, float a4, float a5) {
  return r == 0 ? a0 : r == 1 ? a1 : r == 2 ? a2 : r == 3 ? a3 : r == 4 ? a4 : a5;
}
__device__ __forceinline__ float w_red8(unsigned gm, float v) {
  v += __shfl_xor_sync(gm, v, 1);
  v += __shfl_xor_sync(gm, v, 2);
  v += __shfl_xor_sync(gm, v, 4);
  return v;
}

#define W_INERTIA 1
#define W_FORCE 2
#define W_PB 4
#define W_CONTACTS 8
template <class C>
__device__ __noinline__ void w_inward8(const WModel<C>& M, float* sm, const WLane& w, bool run, int flags, int tmode, int dmode) {
  // fully predicated, warp-uniform control flow: the four 8-lane groups must never diverge from each other, otherwise
  // the hardware serialises them and the row parallelism is lost
  const bool inertia = flags & W_INERTIA, force = flags & W_FORCE;
  const int nlim = (tmode == 0 || dmode == 0) ? ((const int*)sm)[C::lim] : 0;
  const int slot = w.lane >> 3, r = w.lane & 7, gbase = w.lane & ~7;
  const bool rowok = r < 6;
  const int rr = rowok ? r : 5;
  const int* cflag = (const int*)(sm + C::cflag);
  int ridx[6];
#pragma unroll
  for (int j = 0; j < 6; j++) { int i0 = rr < j ? rr : j, j0 = rr < j ? j : rr; ridx[j] = i0 * 6 - (i0 * (i0 - 1)) / 2 + (j0 - i0); }
  for (int t = 0; t < M.sched_T; t++) {
    const int b0 = M.sched[t][slot];
    const bool actv = run && b0 >= 0;
    const int b = actv ? b0 : 0;
    const bool rowact = actv && rowok;
    float Ar[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, pr = 0.f;
    if (inertia && rowact) {
      const float* q = sm + C::irb + 10 * b;
      float m = q[0], cx = q[1], cy = q[2], cz = q[3];
      switch (r) {
        case 0: Ar[0] = q[4]; Ar[1] = q[7]; Ar[2] = q[8]; Ar[4] = -cz; Ar[5] = cy; break;
        case 1: Ar[0] = q[7]; Ar[1] = q[5]; Ar[2] = q[9]; Ar[3] = cz; Ar[5] = -cx; break;
        case 2: Ar[0] = q[8]; Ar[1] = q[9]; Ar[2] = q[6]; Ar[3] = -cy; Ar[4] = cx; break;
        case 3: Ar[1] = cz; Ar[2] = -cy; Ar[3] = m; break;
        case 4: Ar[0] = -cz; Ar[2] = cx; Ar[4] = m; break;
        default: Ar[0] = cy; Ar[1] = -cx; Ar[5] = m; break;
      }
    }
    if (force && (flags & W_PB) && rowact) pr = sm[C::pb + 6 * b + r];
    if (flags & W_CONTACTS) {
      const int hasg = (actv && M.bgeom_adr[b + 1] > M.bgeom_adr[b]) ? 1 : 0;
      const int g = hasg ? M.bgeom_list[M.bgeom_adr[b]] : 0;
      const int s0 = M.slot_adr[g], nsl = hasg ? M.slot_adr[g + 1] - s0 : 0;
      const V3 t1 = ld3(sm + C::ct1 + 3 * g);
      for (int s = 0; s < M.sched_ns[t]; s++) {
        const int c = s < nsl ? s0 + s : s0;
        int fl = s < nsl ? cflag[c] : 0;
        if (!(fl & 1)) fl = 0;
        if (!__any_sync(W_FULL, (fl & 30) != 0)) continue;
        const V3 cp = ld3(sm + C::cpos + 3 * c);
        const float D = sm[C::cD + c];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const bool on = (fl & (2 << k)) != 0;      // select, never multiply: unused slots hold uninitialised shared memory
          S6 xw = w_wrench(M, cp, t1, k);
          if (!on) xw = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
          float xr = rowok ? w_sel6(rr, xw.a.x, xw.a.y, xw.a.z, xw.l.x, xw.l.y, xw.l.z) : 0.f;
          float dx = on ? D * xr : 0.f;
          if (inertia) {
            Ar[0] = fmaf(dx, xw.a.x, Ar[0]); Ar[1] = fmaf(dx, xw.a.y, Ar[1]); Ar[2] = fmaf(dx, xw.a.z, Ar[2]);
            Ar[3] = fmaf(dx, xw.l.x, Ar[3]); Ar[4] = fmaf(dx, xw.l.y, Ar[4]); Ar[5] = fmaf(dx, xw.l.z, Ar[5]);
          }
          if (force && on) pr = fmaf(-dx, sm[C::caref + 4 * c + k], pr);
        }
      }
    }
    {
      const int ca = M.child_adr[b], ncb = actv ? M.child_adr[b + 1] - ca : 0;
      for (int ci = 0; ci < M.sched_nc[t]; ci++) {
        if (ci < ncb && rowok) {
          int c = M.child_list[ca + ci];
          if (inertia) {
#pragma unroll
            for (int j = 0; j < 6; j++) Ar[j] += sm[C::IA + 21 * c + ridx[j]];
          }
          if (force) pr += sm[C::pA + 6 * c + r];
        }
      }
    }
    const int d0 = M.dofadr[b], ndb = actv ? M.dofnum[b] : 0;
    for (int k = M.sched_nd[t] - 1; k >= 0; k--) {
      const bool has = k < ndb;
      const int d = has ? d0 + k : d0;
      S6 S = w_dofS(M, sm, b, has ? k : 0);
      float sr = rowok ? w_sel6(rr, S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z) : 0.f;
      float lD = 0.f, lT = 0.f;
      for (int e = 0; e < nlim; e++) {
        if (((const int*)sm)[C::lim + 4 + 8 * e] == d && (((const int*)sm)[C::lim + 4 + 8 * e + 5] & 1)) {
          lD = WLIM(sm, e, 2); lT = WLIM(sm, e, 1) * lD * WLIM(sm, e, 3);
        }
      }
      float Ur, di;
      if (inertia) {
        Ur = Ar[0] * S.a.x;
        Ur = fmaf(Ar[1], S.a.y, Ur); Ur = fmaf(Ar[2], S.a.z, Ur); Ur = fmaf(Ar[3], S.l.x, Ur); Ur = fmaf(Ar[4], S.l.y, Ur); Ur = fmaf(Ar[5], S.l.z, Ur);
        float D = w_red8(W_FULL, sr * Ur) + M.arm[d] + ((dmode == 1) ? ((d >= 6) ? M.h * M.kd[d - 6] : 0.f) : lD);
        di = has ? 1.0f / D : 0.f;
        float cu = has ? Ur * di : 0.f;
#pragma unroll
        for (int j = 0; j < 6; j++) Ar[j] = fmaf(-cu, __shfl_sync(W_FULL, Ur, gbase + j), Ar[j]);
        if (has && rowok) sm[C::U + 6 * d + r] = Ur;
        if (has && r == 0) sm[C::Dinv + d] = di;
      } else {
        Ur = (has && rowok) ? sm[C::U + 6 * d + r] : 0.f;
        di = has ? sm[C::Dinv + d] : 0.f;
      }
      if (force) {
        float tin = 0.f;
        if (d >= 6) {
          int a = d - 6;
          if (tmode == 0) tin = sm[C::tau + a] + lT;
          else if (tmode == 2) {
            float tgt = fmaf(sm[C::act + a], M.ascale[a], M.aoffset[a]);
            float err = sm[C::qpos + d + 1] + sm[C::qvel + d] * M.h - tgt;
            tin = -M.kp[a] * err - M.kd[a] * sm[C::qvel + d];
          }
        }
        float uu = tin - w_red8(W_FULL, sr * pr);
        if (has && r == 0) sm[C::u + d] = uu;
        if (has) pr = fmaf(Ur, uu * di, pr);
      }
    }
    if (rowact) {
      if (inertia) {
#pragma unroll
        for (int j = 0; j < 6; j++) if (j >= r) sm[C::IA + 21 * b + ridx[j]] = Ar[j];
      }
      if (force) sm[C::pA + 6 * b + r] = pr;
    }
    __syncwarp();
  }
}

template <class C>
__device__ __noinline__ void w_outward8(const WModel<C>& M, float* sm, const WLane& w, bool run, int which, int mode) {
  float* qout = sm + (which == 0 ? C::qacc : which == 1 ? C::qstar : C::spdab);
  const int slot = w.lane >> 3, r = w.lane & 7;
  const bool rowok = r < 6;
  const int rr = rowok ? r : 5;
  for (int t = M.sched_T - 1; t >= 0; t--) {
    const int b0 = M.sched[t][slot];
    const bool actv = run && b0 >= 0;
    const int b = actv ? b0 : 0;
    float ar = (b == 0 || !rowok || !actv) ? 0.f : sm[C::acc + 6 * M.parent[b] + r];
    const int d0 = M.dofadr[b], ndb = actv ? M.dofnum[b] : 0;
    for (int k = 0; k < M.sched_nd[t]; k++) {
      const bool has = k < ndb;
      const int d = has ? d0 + k : d0;
      S6 S = w_dofS(M, sm, b, has ? k : 0);
      float sr = rowok ? w_sel6(rr, S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z) : 0.f;
      float qdd;
      if (mode == 1) qdd = sm[C::qacc + d];
      else {
        float Ur = (has && rowok) ? sm[C::U + 6 * d + r] : 0.f;
        qdd = sm[C::Dinv + d] * (sm[C::u + d] - w_red8(W_FULL, Ur * ar));
        if (has && r == 0) qout[d] = qdd;
      }
      if (has) ar = fmaf(sr, qdd, ar);
    }
    if (actv && rowok) sm[C::acc + 6 * b + r] = ar;
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ABA inward sweep
// flags: 1 INERTIA | 2 FORCE | 4 PB | 8 CONTACTS ; tmode: 0 tau (+limit rows) | 1 zero | 2 stable-PD -kp e - kd qd ; dmode: 0 armature (+limit rows) | 1 + h kd
template <class C>
__device__ __noinline__ void w_inward(const WModel<C>& M, float* sm, const WLane& w, bool run, int flags, int tmode, int dmode, unsigned long long only = 0ull) {
  // only != 0: recompute just the bodies in the mask (solver iterations >= 2: sub-trees without constraint rows keep the
  // articulated inertia / bias force of the first iteration, they do not depend on the working set)
  if (C::LPE == 32 && M.rowpar && M.sched_T > 0) { w_inward8(M, sm, w, run, flags, tmode, dmode); return; }
  const bool inertia = flags & W_INERTIA, force = flags & W_FORCE;
  const int nlim = (tmode == 0 || dmode == 0) ? ((const int*)sm)[C::lim] : 0;
  for (int lev = M.nlevel - 1; lev >= 0; lev--) {
    if (run) {
      for (int i = M.level_adr[lev] + w.li; i < M.level_adr[lev + 1]; i += C::LPE) {
        int b = M.level_list[i];
        if (only && !((only >> b) & 1ull)) continue;
        float A[21];
        S6 p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
        if (inertia) rb_expand(sm + C::irb + 10 * b, A);
        if (force && (flags & W_PB)) p = ld6(sm + C::pb + 6 * b);
        if (flags & W_CONTACTS) {
          const int* cflag = (const int*)(sm + C::cflag);
          for (int gi = M.bgeom_adr[b]; gi < M.bgeom_adr[b + 1]; gi++) {
            int g = M.bgeom_list[gi];
            V3 t1 = ld3(sm + C::ct1 + 3 * g);
            for (int c = M.slot_adr[g]; c < M.slot_adr[g + 1]; c++) {
              int fl = cflag[c];
              if (!(fl & 1) || !(fl & 30)) continue;
              V3 cp = ld3(sm + C::cpos + 3 * c);
              float D = sm[C::cD + c];
              for (int k = 0; k < 4; k++) {
                if (!(fl & (2 << k))) continue;
                S6 xw = w_wrench(M, cp, t1, k);
                float xv[6] = {xw.a.x, xw.a.y, xw.a.z, xw.l.x, xw.l.y, xw.l.z};
                if (inertia) sym_rank1(A, xv, -D);
                if (force) p = p - (D * sm[C::caref + 4 * c + k]) * xw;
              }
            }
          }
        }
        for (int ci = M.child_adr[b]; ci < M.child_adr[b + 1]; ci++) {
          int c = M.child_list[ci];
          if (inertia) {
#pragma unroll
            for (int j = 0; j < 21; j++) A[j] += sm[C::IA + 21 * c + j];
          }
          if (force) p = p + ld6(sm + C::pA + 6 * c);
        }
        int d0 = M.dofadr[b];
        for (int k = M.dofnum[b] - 1; k >= 0; k--) {
          int d = d0 + k;
          S6 S = w_dofS(M, sm, b, k);
          float s[6] = {S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z}, Uv[6], di;
          float lD = 0.f, lT = 0.f;
          for (int e = 0; e < nlim; e++) {   // joint-limit rows of this dof that sit in the working set (rare)
            if (((const int*)sm)[C::lim + 4 + 8 * e] == d && (((const int*)sm)[C::lim + 4 + 8 * e + 5] & 1)) {
              lD = WLIM(sm, e, 2); lT = WLIM(sm, e, 1) * lD * WLIM(sm, e, 3);
            }
          }
          if (inertia) {
            sym_mul(A, s, Uv);
            float D = M.arm[d] + ((dmode == 1) ? ((d >= 6) ? M.h * M.kd[d - 6] : 0.f) : lD);
#pragma unroll
            for (int j = 0; j < 6; j++) D = fmaf(s[j], Uv[j], D);
            di = 1.0f / D;
#pragma unroll
            for (int j = 0; j < 6; j++) sm[C::U + 6 * d + j] = Uv[j];
            sm[C::Dinv + d] = di;
            sym_rank1(A, Uv, di);
          } else {
#pragma unroll
            for (int j = 0; j < 6; j++) Uv[j] = sm[C::U + 6 * d + j];
            di = sm[C::Dinv + d];
          }
          if (force) {
            float tin = 0.f;
            if (d >= 6) {
              int a = d - 6;
              if (tmode == 0) tin = sm[C::tau + a] + lT;
              else if (tmode == 2) {
                float tgt = fmaf(sm[C::act + a], M.ascale[a], M.aoffset[a]);
                float err = sm[C::qpos + d + 1] + sm[C::qvel + d] * M.h - tgt;
                tin = -M.kp[a] * err - M.kd[a] * sm[C::qvel + d];
              }
            }
            float uu = tin - dot6(S, p);
            sm[C::u + d] = uu;
            p = p + (uu * di) * s6(v3(Uv[0], Uv[1], Uv[2]), v3(Uv[3], Uv[4], Uv[5]));
          }
        }
        if (inertia) {
#pragma unroll
          for (int j = 0; j < 21; j++) sm[C::IA + 21 * b + j] = A[j];
        }
        if (force) st6(sm + C::pA + 6 * b, p);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ABA outward sweep.  which: 0 qacc | 1 qstar | 2 spdab ; mode 1: accumulate S*qacc only
template <class C>
__device__ __noinline__ void w_outward(const WModel<C>& M, float* sm, const WLane& w, bool run, int which, int mode) {
  if (C::LPE == 32 && M.rowpar && M.sched_T > 0) { w_outward8(M, sm, w, run, which, mode); return; }
  float* qout = sm + (which == 0 ? C::qacc : which == 1 ? C::qstar : C::spdab);
  for (int lev = 0; lev < M.nlevel; lev++) {
    if (run) {
      for (int i = M.level_adr[lev] + w.li; i < M.level_adr[lev + 1]; i += C::LPE) {
        int b = M.level_list[i];
        S6 a = (b == 0) ? s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)) : ld6(sm + C::acc + 6 * M.parent[b]);
        int d0 = M.dofadr[b], nd = M.dofnum[b];
        for (int k = 0; k < nd; k++) {
          int d = d0 + k;
          S6 S = w_dofS(M, sm, b, k);
          float qdd;
          if (mode == 1) qdd = sm[C::qacc + d];
          else { qdd = sm[C::Dinv + d] * (sm[C::u + d] - dot6(ld6(sm + C::U + 6 * d), a)); qout[d] = qdd; }
          a = a + qdd * S;
        }
        st6(sm + C::acc + 6 * b, a);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ collision (floor vs capsule / sphere / box) + joint-limit rows
template <class C>
__device__ __noinline__ unsigned long long w_collide(const WModel<C>& M, float* sm, const WLane& w, int* nrows_out) {
  unsigned long long mask = 0ull;
  int nrows = 0;
  int* cflag = (int*)(sm + C::cflag);
  int* lim = (int*)(sm + C::lim);
  if (w.live && w.li == 0) lim[0] = 0;
  __syncwarp();
  if (w.live) {
    V3 n = ld3(M.plane_n);
    float h0 = dot(n, ld3(sm + C::qpos) - ld3(M.plane_pos));
    for (int g = w.li; g < M.ng; g += C::LPE) {
      int b = M.gbody[g];
      const float* qq = sm + C::xquat + 4 * b;
      Q4 q; q.w = qq[0]; q.x = qq[1]; q.y = qq[2]; q.z = qq[3];
      float R[9];
      q2mat(q, R);
      V3 c = ld3(sm + C::xpos + 3 * b) + mrot(R, ld3(M.gpos[g]));
      float d0 = h0 + dot(n, c);
      int s0 = M.slot_adr[g], s1 = M.slot_adr[g + 1], cnt = 0;
      const float* gm = M.gmat[g];
      int ty = M.gtype[g];
      V3 t1 = ld3(M.t1_default);
      // pass 1: contact points straight into the slot arrays (cD holds the distance until pass 2); loops kept rolled --
      // this function runs once per substep and its footprint, not its instruction count, is what costs (instruction cache)
      if (ty == SMPLSIM_GEOM_CAPSULE || ty == SMPLSIM_GEOM_SPHERE) {
        V3 axw = mrot(R, v3(gm[2], gm[5], gm[8]));
        float rad = M.gsize[g][0], hl = (ty == SMPLSIM_GEOM_CAPSULE) ? M.gsize[g][1] : 0.f, na = dot(n, axw);
        int nend = (ty == SMPLSIM_GEOM_CAPSULE) ? 2 : 1;
        if (nend == 2) {
          t1 = axw - na * n;
          float nn = sqrtf(dot(t1, t1));
          t1 = (nn < 1e-15f) ? v3(1.f, 0.f, 0.f) : (1.0f / nn) * t1;
        }
#pragma unroll 1
        for (int e = 0; e < nend; e++) {
          float sg = e ? -hl : hl, dist = d0 + sg * na - rad;
          if (dist > M.margin) continue;
          st3(sm + C::cpos + 3 * (s0 + cnt), c + sg * axw - (rad + 0.5f * dist) * n); sm[C::cD + s0 + cnt] = dist; cnt++;
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 8 && cnt < 4; i++) {
          V3 vl = v3((i & 1) ? M.gsize[g][0] : -M.gsize[g][0], (i & 2) ? M.gsize[g][1] : -M.gsize[g][1], (i & 4) ? M.gsize[g][2] : -M.gsize[g][2]);
          V3 wv = mrot(R, mrot(gm, vl));
          float l = dot(n, wv);
          if (d0 + l > M.margin || l > 0.f) continue;
          float dist = d0 + l;
          st3(sm + C::cpos + 3 * (s0 + cnt), c + wv - (0.5f * dist) * n); sm[C::cD + s0 + cnt] = dist; cnt++;
        }
      }
      if (cnt) {
        st3(sm + C::ct1 + 3 * g, t1);
        S6 v = ld6(sm + C::vel + 6 * b);
#pragma unroll 1
        for (int s = 0; s < cnt; s++) {
          int c2 = s0 + s;
          V3 cp = ld3(sm + C::cpos + 3 * c2);
          float pm = sm[C::cD + c2] - M.margin, imp = w_impedance(M, pm);
          float R0 = fmaxf((1.f - imp) / imp * (M.tran_iw0[b] + M.mu * M.mu * M.tran_iw0[b]), 1e-15f);
          float R1 = R0 / fmaxf(M.impratio, 1e-15f), mu = M.mu * sqrtf(R1 / R0);
          sm[C::cD + c2] = 1.0f / (2.f * mu * mu * R0);
          float kterm = M.K * imp * pm;
          int guess = 0;      // rows with aref > 0 (active if the body ends up with ~zero acceleration along the row)
#pragma unroll 1
          for (int k = 0; k < 4; k++) {
            float ar = -M.B * dot6(w_wrench(M, cp, t1, k), v) - kterm;
            sm[C::caref + 4 * c2 + k] = ar;
            if (ar > 0.f) guess |= 2 << k;
          }
          // warmset: start from the working set this slot ended the previous substep with (new contacts: all rows active,
          // or -- bit 3 -- the aref-sign guess)
          int pf = cflag[c2];
          int fresh = (M.warmset & 8) ? guess : ((M.warmset & 2) ? 0 : 30);
          int inh = (M.warmset & 1) ? ((pf & 1) ? (pf & 30) : fresh) : ((M.warmset & 4) ? 30 : 0);
          cflag[c2] = M.warmset ? (1 | inh) : 1;
        }
        mask |= 1ull << (g + 1);
        nrows += 4 * cnt;
      }
#pragma unroll 1
      for (int s = s0 + cnt; s < s1; s++) cflag[s] = 0;
    }
    // joint limits (margin 0): compact list, at most W_MAXLIM rows (more would need >8 joints past +-range at once)
    for (int d = 6 + w.li; d < M.nv; d += C::LPE) {
      if (!M.limited[d]) continue;
      float q = sm[C::qpos + d + 1], dlo = q - M.range[d][0], dhi = M.range[d][1] - q, dist, sg;
      if (dlo < 0.f) { dist = dlo; sg = 1.f; }
      else if (dhi < 0.f) { dist = dhi; sg = -1.f; }
      else continue;
      int e = atomicAdd(&lim[0], 1);
      if (e < W_MAXLIM) {
        float imp = w_impedance(M, dist);
        lim[4 + 8 * e] = d;
        WLIM(sm, e, 1) = sg;
        WLIM(sm, e, 2) = 1.0f / fmaxf((1.f - imp) / imp * M.diw0[d], 1e-15f);
        WLIM(sm, e, 3) = -M.B * sg * sm[C::qvel + d] - M.K * imp * dist;
        WLIM(sm, e, 4) = 0.f;
        lim[4 + 8 * e + 5] = M.warmset ? 1 : 0;
        nrows++;
      }
    }
  }
  __syncwarp();
  if (w.live && w.li == 0 && lim[0] > W_MAXLIM) lim[0] = W_MAXLIM;
  __syncwarp();
  *nrows_out = nrows;
  return mask;
}

// ------------------------------------------------------------------ constraint rows
// mode 0: working set := (r < 0) at acc / qacc (warm start) ; mode 1: compare (r* < 0) at acc / qstar with the working set
template <class C>
__device__ __noinline__ bool w_eval_rows(const WModel<C>& M, float* sm, const WLane& w, bool run, int mode) {
  bool same = true;
  if (run) {
    int* cflag = (int*)(sm + C::cflag);
    for (int c = w.li; c < M.nslot; c += C::LPE) {
      int fl = cflag[c];
      if (!(fl & 1)) continue;
      int g = M.slot_geom[c], b = M.gbody[g];
      S6 a = ld6(sm + C::acc + 6 * b);
      V3 cp = ld3(sm + C::cpos + 3 * c), t1 = ld3(sm + C::ct1 + 3 * g);
      int nf = 1;
      for (int k = 0; k < 4; k++)
        if (dot6(w_wrench(M, cp, t1, k), a) - sm[C::caref + 4 * c + k] < 0.f) nf |= 2 << k;
      if (mode == 0) cflag[c] = nf;
      else if (nf != fl) same = false;
    }
    int nlim = ((int*)sm)[C::lim];
    for (int e = w.li; e < nlim; e += C::LPE) {
      int d = ((int*)sm)[C::lim + 4 + 8 * e];
      float r = WLIM(sm, e, 1) * sm[(mode == 0 ? C::qacc : C::qstar) + d] - WLIM(sm, e, 3);
      int nf = r < 0.f ? 1 : 0, fl = ((int*)sm)[C::lim + 4 + 8 * e + 5];
      if (mode == 0) ((int*)sm)[C::lim + 4 + 8 * e + 5] = nf;
      else if (nf != (fl & 1)) same = false;
    }
  }
  bool r = w_gall(same || !run, w);
  __syncwarp();
  return r;
}

// row-space pass over this lane's rows.  r = x.acc2 - aref (current iterate), rs = x.acc - aref (trial), d = rs - r.
// op 0: adopt  (phi := -D rs on the used set ; set := rs<0)                      [acc2 := acc done by the caller]
// op 1: sums   (g1 += d phi ; g2 += d (phis - phi) ; s1,s2 at step al)
// op 2: apply  (phi += al (phis - phi) ; set := (r + al d < 0))
template <class C>
__device__ __noinline__ void w_rows(const WModel<C>& M, float* sm, const WLane& w, bool run, int op, float al, float* out4) {
  float g1 = 0.f, g2 = 0.f, s1 = 0.f, s2 = 0.f;
  if (run) {
    int* cflag = (int*)(sm + C::cflag);
    for (int c = w.li; c < M.nslot; c += C::LPE) {
      int fl = cflag[c];
      if (!(fl & 1)) continue;
      int g = M.slot_geom[c], b = M.gbody[g];
      S6 a1 = ld6(sm + C::acc + 6 * b), a2 = ld6(sm + C::acc2 + 6 * b);
      V3 cp = ld3(sm + C::cpos + 3 * c), t1 = ld3(sm + C::ct1 + 3 * g);
      float D = sm[C::cD + c];
      int nf = 1;
      for (int k = 0; k < 4; k++) {
        S6 xw = w_wrench(M, cp, t1, k);
        float ar = sm[C::caref + 4 * c + k], rs = dot6(xw, a1) - ar;
        float phs = (fl & (2 << k)) ? -D * rs : 0.f;
        if (op == 0) { sm[C::cphi + 4 * c + k] = phs; if (rs < 0.f) nf |= 2 << k; }
        else {
          float r = dot6(xw, a2) - ar, d = rs - r, ph = sm[C::cphi + 4 * c + k], v = fmaf(al, d, r);
          if (op == 1) {
            g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
            if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
          } else { sm[C::cphi + 4 * c + k] = fmaf(al, phs - ph, ph); if (v < 0.f) nf |= 2 << k; }
        }
      }
      if (op != 1) cflag[c] = nf;
    }
    int nlim = ((int*)sm)[C::lim];
    for (int e = w.li; e < nlim; e += C::LPE) {
      int d_ = ((int*)sm)[C::lim + 4 + 8 * e], fl = ((int*)sm)[C::lim + 4 + 8 * e + 5];
      float sg = WLIM(sm, e, 1), D = WLIM(sm, e, 2), ar = WLIM(sm, e, 3);
      float rs = sg * sm[C::qstar + d_] - ar, phs = (fl & 1) ? -D * rs : 0.f;
      if (op == 0) { WLIM(sm, e, 4) = phs; ((int*)sm)[C::lim + 4 + 8 * e + 5] = rs < 0.f ? 1 : 0; }
      else {
        float r = sg * sm[C::qacc + d_] - ar, d = rs - r, ph = WLIM(sm, e, 4), v = fmaf(al, d, r);
        if (op == 1) {
          g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
          if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
        } else { WLIM(sm, e, 4) = fmaf(al, phs - ph, ph); ((int*)sm)[C::lim + 4 + 8 * e + 5] = v < 0.f ? 1 : 0; }
      }
    }
  }
  out4[0] = g1; out4[1] = g2; out4[2] = s1; out4[3] = s2;
}

// bodies whose articulated quantities depend on the working set: those carrying contact slots or limit rows, and their ancestors
template <class C>
__device__ __noinline__ unsigned long long w_dirty(const WModel<C>& M, const float* sm, const WLane& w, unsigned long long geom_mask) {
  // own bit per body (lane-parallel, ballot), then ancestor closure (bodies are in topological order: parent < child)
  const int nlim = w.live ? ((const int*)sm)[C::lim] : 0;
  unsigned long long dm = 1ull;
#pragma unroll 1
  for (int b0 = 0; b0 < M.nb; b0 += C::LPE) {
    int b = b0 + w.li;
    bool d = false;
    if (w.live && b > 0 && b < M.nb) {
      for (int gi = M.bgeom_adr[b]; gi < M.bgeom_adr[b + 1]; gi++) d |= ((geom_mask >> (M.bgeom_list[gi] + 1)) & 1ull) != 0ull;
      for (int e = 0; e < nlim; e++) { int dd = ((const int*)sm)[C::lim + 4 + 8 * e] - M.dofadr[b]; d |= (dd >= 0 && dd < M.dofnum[b]); }
    }
    unsigned bal = (__ballot_sync(W_FULL, d) & w.gmask) >> (w.lane - w.li);
    dm |= (unsigned long long)bal << b0;
  }
#pragma unroll 1
  for (int b = M.nb - 1; b > 0; b--)
    if ((dm >> b) & 1ull) dm |= 1ull << M.parent[b];
  return dm;
}

#ifdef SMPLSIM_TRACE
__device__ float g_trace[8192];
__device__ int g_trace_n;
#define W_TRACE(...) do { if (w.live && w.li == 0) { float tv_[] = {__VA_ARGS__}; int n_ = sizeof(tv_) / 4, o_ = atomicAdd(&g_trace_n, n_ + 1); \
  if (o_ + n_ + 1 <= 8192) { g_trace[o_] = (float)n_; for (int i_ = 0; i_ < n_; i_++) g_trace[o_ + 1 + i_] = tv_[i_]; } } } while (0)
#else
#define W_TRACE(...)
#endif
// ------------------------------------------------------------------ constraint solve (active-set Newton, each system one ABA); returns extra solves
template <class C>
__device__ __noinline__ int w_solve(const WModel<C>& M, float* sm, const WLane& w, bool any_rows, unsigned long long geom_mask) {
  bool plain = w.live && !any_rows;
  if (__any_sync(W_FULL, plain)) {
    w_inward(M, sm, w, plain, W_INERTIA | W_FORCE | W_PB, 0, 0);
    w_outward(M, sm, w, plain, 0, 0);
  }
  bool run = w.live && any_rows;
  if (!__any_sync(W_FULL, run)) return 0;
  if (!M.warmset) {            // initial working set from the previous qacc (MuJoCo-style warm start) ...
    w_outward(M, sm, w, run, 0, 1);
    w_eval_rows(M, sm, w, run, 0);
  }                            // ... or inherited per contact slot from the previous substep (set in w_collide)
  bool have_point = false;
  int it = 0, iters = 0;
  float o4[4];
  unsigned long long dirty = 0ull;
  for (; it < W_SOLVER_MAXITER; it++) {
    if (!__any_sync(W_FULL, run)) break;
    if (it == 1 && M.dirtypath) dirty = w_dirty(M, sm, w, geom_mask);
    w_inward(M, sm, w, run, W_INERTIA | W_FORCE | W_PB | W_CONTACTS, 0, 0, dirty);
    w_outward(M, sm, w, run, 1, 0);
    bool same = w_eval_rows(M, sm, w, run, 1);
    bool fin = run && same, adopt = run && !same && !have_point, lsrch = run && !same && have_point;
#ifdef SMPLSIM_TRACE
    { float mx = 0.f; for (int d = 0; d < M.nv; d++) mx = fmaxf(mx, fabsf(sm[C::qstar + d]));
      int* cf = (int*)(sm + C::cflag); float code = 0.f; for (int c = 0; c < M.nslot; c++) if (cf[c] & 1) code += 1.f;
      W_TRACE(1.f, (float)it, fin ? 1.f : (adopt ? 2.f : (lsrch ? 3.f : 0.f)), mx, code); }
#endif
    if (__any_sync(W_FULL, lsrch)) {   // exact line search between the iterate (qacc, acc2) and the trial point (qstar, acc), row space only
      w_rows(M, sm, w, lsrch, 1, 0.f, o4);
      float g1 = w_gsum<C>(o4[0]), g2 = w_gsum<C>(o4[1]), s1 = w_gsum<C>(o4[2]), s2;
      float f0 = g1 + s1, al = 0.f, lo = 0.f, hi = -1.f, tol = M.ls_tol * fabsf(f0);
      bool searching = lsrch && (f0 < -W_LS_NOISE * (fabsf(g1) + fabsf(s1)));   // |f0| below the fp32 cancellation floor: converged
      if (searching) al = 1.f;
      for (int ls = 0; ls < W_LS_MAXITER; ls++) {
        if (!__any_sync(W_FULL, searching)) break;
        w_rows(M, sm, w, searching, 1, al, o4);
        s1 = w_gsum<C>(o4[2]); s2 = w_gsum<C>(o4[3]);
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
      W_TRACE(2.f, f0, al, g1, g2);
      if (lsrch && !step) { run = false; iters = it; }
      w_rows(M, sm, w, step, 2, al, o4);
      __syncwarp();
      if (step) {
        for (int d = w.li; d < M.nv; d += C::LPE) sm[C::qacc + d] = fmaf(al, sm[C::qstar + d] - sm[C::qacc + d], sm[C::qacc + d]);
        for (int j = w.li; j < 6 * M.nb; j += C::LPE) sm[C::acc2 + j] = fmaf(al, sm[C::acc + j] - sm[C::acc2 + j], sm[C::acc2 + j]);
      }
    }
    if (adopt) w_rows(M, sm, w, true, 0, 0.f, o4);
    __syncwarp();   // w_rows loads acc2 / qstar (lane per slot) before the lane-per-element copies below overwrite them
    if (fin || adopt) {
      for (int d = w.li; d < M.nv; d += C::LPE) sm[C::qacc + d] = sm[C::qstar + d];
      if (adopt) for (int j = w.li; j < 6 * M.nb; j += C::LPE) sm[C::acc2 + j] = sm[C::acc + j];
    }
    if (fin) { run = false; iters = it; }
    if (adopt) have_point = true;
    __syncwarp();
  }
  if (run) iters = it;
  return iters;
}

// ------------------------------------------------------------------ stable PD   (controllers.py:116-190)
// ONE ABA per substep: a = (M_s + h Kd)^-1 (-C_s - Kp e - Kd v) with the inertia / bias of the state the FK arrays describe
// (s = state of the last forward pass, quirk Q1) and e, v of the state currently in qpos / qvel.  Callers run it right after
// the integration of substep k (FK arrays still at s_k, qpos / qvel already at s_{k+1}), so the torque of substep k+1 is an
// elementwise formula -- the separate "force-only" solve of the first version is gone.
template <class C>
__device__ __noinline__ void w_spd_prepare(const WModel<C>& M, float* sm, const WLane& w) {
  w_inward(M, sm, w, w.live, W_INERTIA | W_FORCE | W_PB, 2, 1);
  w_outward(M, sm, w, w.live, 2, 0);
}

template <class C>
__device__ __noinline__ void w_torque(const WModel<C>& M, float* sm, const WLane& w, const SmplsimState& st, int env) {
  int mode = M.cfg.control_mode;
  if (w.live) {
    for (int i = w.li; i < M.nu; i += C::LPE) {
      float a = sm[C::act + i], tq;
      if (mode == SMPLSIM_CTRL_TORQUE) tq = a * M.ascale[i];
      else if (mode == SMPLSIM_CTRL_SIMPLE_PID) {   // stateful: integral / last error live in HBM (L2-resident, 2 x nu words per env)
        size_t o = (size_t)env * M.nu + i;
        float dt = M.h * (float)M.cfg.nsubsteps, lim = M.tlim[i];
        float err = fmaf(a, M.ascale[i], M.aoffset[i]) - sm[C::qpos + 7 + i], le = st.pid_last_error[o];
        float derr = (le != le) ? 0.f : err - le;
        float in = fminf(fmaxf(fmaf(err, dt, st.pid_integral[o]), -lim), lim);
        st.pid_integral[o] = in; st.pid_last_error[o] = err;
        tq = M.kp[i] * err + in + M.kd[i] * derr / dt;
      } else {
        float tgt = fmaf(a, M.ascale[i], M.aoffset[i]), q = sm[C::qpos + 7 + i], qd = sm[C::qvel + 6 + i];
        if (mode == SMPLSIM_CTRL_PD) tq = -M.kp[i] * (q - tgt) - M.kd[i] * qd;
        else tq = -M.kp[i] * (q + qd * M.h - tgt) - M.kd[i] * (qd + sm[C::spdab + 6 + i] * M.h);
      }
      sm[C::tau + i] = fminf(fmaxf(tq, -M.tlim[i]), M.tlim[i]);
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------ semi-implicit Euler; lanes li = 0..2 return the root displacement component
template <class C>
__device__ __noinline__ float w_integrate(const WModel<C>& M, float* sm, const WLane& w) {
  float disp = 0.f;
  if (w.live) {
    float h = M.h;
    for (int d = w.li; d < M.nv; d += C::LPE) {
      float v = fmaf(h, sm[C::qacc + d], sm[C::qvel + d]);
      sm[C::qvel + d] = v;
      if (d < 3) { disp = h * v; sm[C::qpos + d] += disp; }
      else if (d >= 6) sm[C::qpos + d + 1] = fmaf(h, v, sm[C::qpos + d + 1]);
    }
  }
  __syncwarp();
  if (w.live && w.li == 0) {
    float* qpos = sm + C::qpos;
    V3 wv = ld3(sm + C::qvel + 3);
    float n = sqrtf(dot(wv, wv)), ang = n * M.h;
    Q4 q; q.w = qpos[3]; q.x = qpos[4]; q.y = qpos[5]; q.z = qpos[6];
    if (ang > 0.f) {
      float sn, cs;
      w_sincos(0.5f * ang, &sn, &cs);
      float s = sn / n;
      Q4 dq; dq.w = cs; dq.x = wv.x * s; dq.y = wv.y * s; dq.z = wv.z * s;
      q = qmul(q, dq);
    }
    q = qnormalize(q);
    qpos[3] = q.w; qpos[4] = q.x; qpos[5] = q.y; qpos[6] = q.z;
  }
  __syncwarp();
  return disp;
}

// ====================================================================================================================
// env-level kernels (v3)
// ====================================================================================================================
#define W_TSK_CHANGE 4
#define W_TSK_CURT 5
#define W_TSK_RECOV 6
#define W_TSK_RNG 7

struct WStepArgs {
  SmplsimState st;
  SmplsimAux aux;
  const float* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  int n, nsub, mode;
  int align;   // P > 0: CTA barrier every P substeps (phase alignment for the instruction cache); 0: none
};
struct WResetArgs {
  SmplsimState st;
  SmplsimAux aux;
  const uint8_t* mask;
  const float* qpos0;
  const float* qvel0;
  float* obs;
  int n, init_mode;
};

template <class C>
__device__ __forceinline__ void w_copy(float* dst, const float* src, int n, const WLane& w) {
  if (w.live) for (int i = w.li; i < n; i += C::LPE) dst[i] = src[i];
}

template <class C>
__device__ __noinline__ void w_task_io(const WModel<C>& M, float* sm, const WLane& w, const SmplsimState& st, int env, bool store) {
  if (store) __syncwarp();
  if (w.live && w.li == 0) {
    float* t = sm + C::tsk;
    int* ti = (int*)t;
    if (!store) {
      for (int j = 0; j < 4; j++) t[j] = st.task_target[4 * env + j];
      ti[W_TSK_CHANGE] = st.task_change_step[env]; ti[W_TSK_CURT] = st.progress[env]; ti[W_TSK_RECOV] = st.recovery[env];
      ti[W_TSK_RNG] = (int)st.rng_counter[env];
    } else {
      for (int j = 0; j < 4; j++) st.task_target[4 * env + j] = t[j];
      st.task_change_step[env] = ti[W_TSK_CHANGE]; st.progress[env] = ti[W_TSK_CURT]; st.recovery[env] = ti[W_TSK_RECOV];
      st.rng_counter[env] = (uint32_t)ti[W_TSK_RNG];
    }
  }
  if (!store) __syncwarp();
}

template <class C>
__device__ __noinline__ void w_reset_task(const WModel<C>& M, float* sm, int env) {   // lane li == 0 only
  const SmplsimEnvCfg& c = M.cfg;
  if (c.task == SMPLSIM_TASK_NONE) return;
  float* t = sm + C::tsk;
  int* ti = (int*)t;
  uint32_t r[4];
  philox4x32((uint32_t)ti[W_TSK_RNG], (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
  ti[W_TSK_RNG] = ti[W_TSK_RNG] + 1;
  if (c.task == SMPLSIM_TASK_SPEED) t[0] = (float)(c.tar_speed_max - c.tar_speed_min) * u01(r[0]) + (float)c.tar_speed_min;
  else if (c.task == SMPLSIM_TASK_REACH) {
    t[0] = (float)c.tar_dist_max * (2.0f * u01(r[0]) - 1.0f);
    t[1] = (float)c.tar_dist_max * (2.0f * u01(r[1]) - 1.0f);
    t[2] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[2]) + (float)c.tar_height_min;
  } else t[0] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[0]) + (float)c.tar_height_min;
  ti[W_TSK_CHANGE] = ti[W_TSK_CURT] + rand_range(r[3], c.change_steps_min, c.change_steps_max);
}

// mj_checkPos / mj_checkVel (what 0, before the forward pass) and mj_checkAcc (what 1, after the solve), SURVEY A.2
// ([MJ-upstream] engine_forward.c): a NaN or |x| > mjMAXVAL = 1e10 raises the warning bit and auto-resets the env's data like
// mj_resetData (qpos = qpos0, qvel = ctrl = qacc_warmstart = 0).  Returns the warning bits of this lane's env (1 qpos | 2 qvel | 4 qacc).
#define W_MAXVAL 1e10f
template <class C>
__device__ __noinline__ int w_check(const WModel<C>& M, float* sm, const WLane& w, int what) {
  bool b0 = false, b1 = false;
  if (w.live) {
    if (what == 0) {
#pragma unroll 1
      for (int i = w.li; i < M.nv + 1; i += C::LPE) b0 |= !(fabsf(sm[C::qpos + i]) <= W_MAXVAL);
#pragma unroll 1
      for (int i = w.li; i < M.nv; i += C::LPE) b1 |= !(fabsf(sm[C::qvel + i]) <= W_MAXVAL);
    } else {
#pragma unroll 1
      for (int i = w.li; i < M.nv; i += C::LPE) b0 |= !(fabsf(sm[C::qacc + i]) <= W_MAXVAL);
    }
  }
  b0 = w_gany(b0, w); b1 = w_gany(b1, w);
  int bits = (what == 0) ? (b0 ? 1 : (b1 ? 2 : 0)) : (b0 ? 4 : 0);
  if (bits && w.live) {
#pragma unroll 1
    for (int i = w.li; i < M.nv + 1; i += C::LPE) sm[C::qpos + i] = (i < 3) ? M.bpos[0][i] : (i < 7) ? M.bquat[0][i - 3] : 0.f;
#pragma unroll 1
    for (int i = w.li; i < M.nv; i += C::LPE) { sm[C::qvel + i] = 0.f; sm[C::qacc + i] = 0.f; }
#pragma unroll 1
    for (int i = w.li; i < M.nu; i += C::LPE) sm[C::tau + i] = 0.f;
  }
  __syncwarp();
  return bits;
}

struct WFwd { unsigned long long mask; int iters; int status; };

template <class C>
__device__ __noinline__ float w_substeps(const WModel<C>& M, float* sm, const WLane& w, int nsub, int raw, WFwd* fo, const SmplsimState& st, int env,
                                         bool write_fwd, bool prep_last, int align) {
  const bool spd = (M.cfg.control_mode == SMPLSIM_CTRL_UHC_PD), stale = M.cfg.spd_stale != 0;
  float disp = 0.f;
  bool restore = false;   // raw mode: the caller's ctrl (kept in act) comes back the substep after an auto-reset zeroed it
  for (int s = 0; s < nsub; s++) {
    // keep the warps of a CTA in the same phase: the hot code of one phase fits the instruction cache, that of
    // 14 drifting warps does not (round-1 profile: "no_instruction" was the top stall)
    if ((align & 255) > 0 && (s % (align & 255)) == 0) {
      const int G = align >> 8;              // warps per barrier group (0: the whole CTA)
      if (G == 0) __syncthreads();
      else {
        int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5, grp = wib / G, cnt = min(G, wpb - grp * G) * 32;
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(cnt) : "memory");
      }
    }
    bool did_fk = false;
    if (!raw) {
      if (spd && !stale) { w_fk(M, sm, w, true); w_spd_prepare(M, sm, w); did_fk = true; }
      w_torque(M, sm, w, st, env);
    } else if (__any_sync(W_FULL, restore)) {
      if (w.live && restore) for (int i = w.li; i < M.nu; i += C::LPE) sm[C::tau + i] = sm[C::act + i];
      restore = false;
      __syncwarp();
    }
    int bad = w_check(M, sm, w, 0);
    if (__any_sync(W_FULL, bad != 0)) did_fk = false;
    if (!did_fk) w_fk(M, sm, w, true);
    int nrows = 0;
    unsigned long long m = w_collide(M, sm, w, &nrows);
    unsigned lo = w_gor<C>((unsigned)(m & 0xffffffffull)), hi = w_gor<C>((unsigned)(m >> 32));
    fo->mask = ((unsigned long long)hi << 32) | lo;
    bool any_rows = w_gany(w.live && nrows > 0, w);
    fo->iters = w_solve(M, sm, w, any_rows, fo->mask);
    int badacc = w_check(M, sm, w, 1);
    if (__any_sync(W_FULL, badacc != 0)) {   // mj_checkAcc: forward pass again on the reset data, then integrate
      WLane w2 = w;
      w2.live = w.live && badacc != 0;
      w_fk(M, sm, w2, true);
      int nrows2 = 0;
      unsigned long long m2 = w_collide(M, sm, w2, &nrows2);
      unsigned lo2 = w_gor<C>((unsigned)(m2 & 0xffffffffull)), hi2 = w_gor<C>((unsigned)(m2 >> 32));
      bool any2 = w_gany(w2.live && nrows2 > 0, w2);
      int it2 = w_solve(M, sm, w2, any2, ((unsigned long long)hi2 << 32) | lo2);
      if (badacc) { fo->mask = ((unsigned long long)hi2 << 32) | lo2; fo->iters = it2; }
    }
    bad |= badacc;
    fo->status |= bad;
    if (raw && bad) restore = true;
    if (s == nsub - 1) {
      if (w.live) {
        for (int b = w.li; b < M.nb; b += C::LPE) {   // framelinvel / frameangvel of the last forward pass (quirk Q2), parked in acc2
          S6 v = ld6(sm + C::vel + 6 * b);
          st3(sm + C::acc2 + 6 * b, v.l + cross(v.a, ld3(sm + C::xpos + 3 * b)));
          st3(sm + C::acc2 + 6 * b + 3, v.a);
        }
      }
      if (write_fwd) {
        w_copy<C>(st.qpos_fwd + (size_t)env * (M.nv + 1), sm + C::qpos, M.nv + 1, w);
        w_copy<C>(st.qvel_fwd + (size_t)env * M.nv, sm + C::qvel, M.nv, w);
      }
    }
    disp += w_integrate(M, sm, w);
    if (spd && stale && !raw && (s < nsub - 1 || prep_last)) w_spd_prepare(M, sm, w);   // FK arrays: s_k ; qpos/qvel: s_{k+1}
  }
  return disp;
}

template <class C>
__device__ __forceinline__ Q4 w_heading_inv(const WModel<C>& M, Q4 root) {
  if (!M.cfg.upright_start) { Q4 bc; bc.w = 0.5f; bc.x = -0.5f; bc.y = -0.5f; bc.z = -0.5f; root = qmul(root, bc); }
  V3 rd = qrot_ref(root, v3(1.f, 0.f, 0.f));
  float hd = atan2f(rd.y, rd.x), sn, cs;
  w_sincos(-0.5f * hd, &sn, &cs);
  Q4 h; h.w = cs; h.x = 0.f; h.y = 0.f; h.z = sn;
  return qnormalize(h);
}

// compute_observations: staged in shared memory (aliases the IA accumulators), then streamed out coalesced
template <class C>
__device__ __noinline__ void w_write_obs(const WModel<C>& M, float* sm, const WLane& w, float* obs_row) {
  if (w.live) {
    float* ob = sm + C::obs;
    const float *qpos = sm + C::qpos, *xq = sm + C::xquat, *qvel = sm + C::qvel, *sens = sm + C::acc2;
    int nb = M.nb;
    Q4 r0; r0.w = xq[0]; r0.x = xq[1]; r0.y = xq[2]; r0.z = xq[3];
    Q4 hq = w_heading_inv(M, r0);
    int o = M.cfg.root_height_obs ? 1 : 0, o_rot = o + 3 * (nb - 1), o_vel = o_rot + 6 * nb;
    if (o && w.li == 0) ob[0] = qpos[2];
    for (int b = w.li; b < nb; b += C::LPE) {
      if (b > 0) st3(ob + o + 3 * (b - 1), qrot_ref(hq, ld3(sm + C::xpos + 3 * b)));
      Q4 q; q.w = xq[4 * b]; q.x = xq[4 * b + 1]; q.y = xq[4 * b + 2]; q.z = xq[4 * b + 3];
      Q4 lq = qmul(hq, q);
      st3(ob + o_rot + 6 * b, qrot_ref(lq, v3(1.f, 0.f, 0.f)));
      st3(ob + o_rot + 6 * b + 3, qrot_ref(lq, v3(0.f, 0.f, 1.f)));
      if (M.cfg.self_obs_v == 2) {
        st3(ob + o_vel + 3 * b, qrot_ref(hq, ld3(sens + 6 * b)));
        st3(ob + o_vel + 3 * nb + 3 * b, qrot_ref(hq, ld3(sens + 6 * b + 3)));
      }
    }
    if (M.cfg.self_obs_v == 1) {
      if (w.li == 0) st3(ob + o_vel, qrot_ref(hq, ld3(qvel)));
      if (w.li == 1) st3(ob + o_vel + 3, qrot_ref(hq, ld3(qvel + 3)));
      for (int i = w.li; i < M.nu; i += C::LPE) ob[o_vel + 6 + i] = qvel[6 + i];
    }
    if (w.li == 2) {
      const float* t = sm + C::tsk;
      int ot = M.self_obs_dim;
      Q4 rq; rq.w = qpos[3]; rq.x = qpos[4]; rq.y = qpos[5]; rq.z = qpos[6];
      if (M.cfg.task == SMPLSIM_TASK_SPEED) {
        V3 d = qrot_ref(w_heading_inv(M, rq), v3(1.f, 0.f, 0.f));
        ob[ot] = d.x; ob[ot + 1] = d.y; ob[ot + 2] = t[0];
      } else if (M.cfg.task == SMPLSIM_TASK_REACH) st3(ob + ot, qrot_ref(w_heading_inv(M, rq), ld3(t) - ld3(qpos)));
      else if (M.cfg.task == SMPLSIM_TASK_GETUP) ob[ot] = t[0];
    }
  }
  __syncwarp();
  if (w.live && obs_row) for (int i = w.li; i < M.obs_dim; i += C::LPE) obs_row[i] = sm[C::obs + i];
  __syncwarp();
}

template <class C>
__device__ __noinline__ void w_write_aux(const WModel<C>& M, float* sm, const WLane& w, const SmplsimAux& aux, int env, const WFwd& fo) {
  if (!w.live) return;
  V3 root = ld3(sm + C::qpos);
  int nb = M.nb;
  for (int b = w.li; b < nb; b += C::LPE) {
    size_t bi = (size_t)env * nb + b;
    if (aux.xpos) st3(aux.xpos + bi * 3, ld3(sm + C::xpos + 3 * b) + root);
    if (aux.body_linvel) st3(aux.body_linvel + bi * 3, ld3(sm + C::acc2 + 6 * b));
    if (aux.body_angvel) st3(aux.body_angvel + bi * 3, ld3(sm + C::acc2 + 6 * b + 3));
  }
  if (aux.xquat) w_copy<C>(aux.xquat + (size_t)env * nb * 4, sm + C::xquat, 4 * nb, w);
  if (aux.qacc) w_copy<C>(aux.qacc + (size_t)env * M.nv, sm + C::qacc, M.nv, w);
  if (aux.ctrl) w_copy<C>(aux.ctrl + (size_t)env * M.nu, sm + C::tau, M.nu, w);
  if (w.li == 0) {
    if (aux.contact_mask) aux.contact_mask[env] = fo.mask;
    if (aux.solver_iter) aux.solver_iter[env] = fo.iters;
    if (aux.status) aux.status[env] = (uint8_t)fo.status;
  }
}

extern __shared__ float w_smem[];

template <class C>
__device__ __forceinline__ float* w_setup(const DevModel* G, const WModel<C>*& Mp, WLane& w, int& env, int n) {
  WModel<C>* M = (WModel<C>*)w_smem;
  w_stage_model<C>(G, *M);
  __syncthreads();
  Mp = M;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  int sub = lane / C::LPE;
  w.lane = lane; w.li = lane % C::LPE;
  w.gmask = (C::LPE == 32) ? 0xffffffffu : (((1u << C::LPE) - 1u) << (sub * C::LPE));
  env = (blockIdx.x * wpb + wib) * C::EPW + sub;
  w.live = env < n;
  constexpr int mwords = (int)((sizeof(WModel<C>) + 15) / 16) * 4;
  return w_smem + mwords + (size_t)(wib * C::EPW + sub) * C::total;
}

template <class C>
__global__ void __launch_bounds__(512) k_step3(const DevModel* __restrict__ G, WStepArgs a) {
  const WModel<C>* Mp; WLane w; int env;
  float* sm = w_setup<C>(G, Mp, w, env, a.n);
  const WModel<C>& M = *Mp;
  size_t eo = w.live ? (size_t)env : 0;   // warps without a live env keep running (predicated) so CTA barriers stay legal
  const bool spd = (M.cfg.control_mode == SMPLSIM_CTRL_UHC_PD);
  if (w.live && w.li == 0) ((int*)sm)[C::lim] = 0;
  if (w.live) for (int c = w.li; c < M.nslot; c += C::LPE) ((int*)sm)[C::cflag + c] = 0;   // no inherited working set at launch
  if (spd && M.cfg.spd_stale && a.mode == 0) {   // factors of (M + h Kd) at the state of the last forward pass (quirk Q1)
    w_copy<C>(sm + C::qpos, a.st.qpos_fwd + eo * (M.nv + 1), M.nv + 1, w);
    w_copy<C>(sm + C::qvel, a.st.qvel_fwd + eo * M.nv, M.nv, w);
    __syncwarp();
    w_fk(M, sm, w, true);
  }
  w_copy<C>(sm + C::qpos, a.st.qpos + eo * (M.nv + 1), M.nv + 1, w);
  w_copy<C>(sm + C::qvel, a.st.qvel + eo * M.nv, M.nv, w);
  w_copy<C>(sm + C::qacc, a.st.qacc_warm + eo * M.nv, M.nv, w);
  w_copy<C>(sm + C::act, a.action + eo * M.nu, M.nu, w);
  if (a.mode != 0) w_copy<C>(sm + C::tau, a.action + eo * M.nu, M.nu, w);   // raw ctrl (act keeps a copy for the substep after an auto-reset)
  __syncwarp();
  if (spd && M.cfg.spd_stale && a.mode == 0) w_spd_prepare(M, sm, w);
  w_task_io(M, sm, w, a.st, env, false);
  if (a.mode == 0 && w.live && w.li == 0) {
    int* ti = (int*)(sm + C::tsk);
    if (M.cfg.task != SMPLSIM_TASK_NONE && ti[W_TSK_CURT] >= ti[W_TSK_CHANGE]) w_reset_task(M, sm, env);
  }
  __syncwarp();
  WFwd fo; fo.mask = 0ull; fo.iters = 0; fo.status = 0;
  float disp = w_substeps(M, sm, w, a.nsub, a.mode, &fo, a.st, env, true, false, a.align);
  w_fk(M, sm, w, false);
  if (a.mode == 0) {
    int* ti = (int*)(sm + C::tsk);
    if (w.live && w.li == 0) ti[W_TSK_CURT] += 1;
    __syncwarp();
    w_write_obs(M, sm, w, a.obs ? a.obs + eo * M.obs_dim : nullptr);
    int base = w.lane - w.li;
    float dx = __shfl_sync(W_FULL, disp, base), dy = __shfl_sync(W_FULL, disp, base + 1);
    if (w.live && w.li == 0) {
      const SmplsimEnvCfg& c = M.cfg;
      const float* t = sm + C::tsk;
      float rew = 0.f;
      if (c.task == SMPLSIM_TASK_SPEED) {
        float inv_dt = 1.0f / (M.h * (float)a.nsub), vx = dx * inv_dt, vy = dy * inv_dt, e = t[0] - vx;
        rew = expf(-0.25f * (e * e + 0.1f * vy * vy));
      } else if (c.task == SMPLSIM_TASK_REACH) {
        V3 dl = ld3(t) - (ld3(sm + C::xpos + 3 * c.reach_body) + ld3(sm + C::qpos));
        rew = expf(-4.0f * dot(dl, dl));
      } else if (c.task == SMPLSIM_TASK_GETUP) { float e = t[0] - sm[C::qpos + 2]; rew = expf(-4.0f * e * e); }
      int term = 0, trunc = 0, pass_time = ti[W_TSK_CURT] > c.episode_length;
      if (c.task == SMPLSIM_TASK_NONE) trunc = pass_time;
      else if (c.task == SMPLSIM_TASK_GETUP && ti[W_TSK_RECOV] > 0) ti[W_TSK_RECOV] -= 1;
      else { trunc = pass_time; term = (fo.mask & ~M.legal_mask) != 0ull; }
      if (a.reward) a.reward[env] = rew;
      if (a.terminated) a.terminated[env] = (uint8_t)term;
      if (a.truncated) a.truncated[env] = (uint8_t)trunc;
    }
  }
  w_write_aux(M, sm, w, a.aux, env, fo);
  w_copy<C>(a.st.qpos + eo * (M.nv + 1), sm + C::qpos, M.nv + 1, w);
  w_copy<C>(a.st.qvel + eo * M.nv, sm + C::qvel, M.nv, w);
  w_copy<C>(a.st.qacc_warm + eo * M.nv, sm + C::qacc, M.nv, w);
  if (a.mode == 0) w_task_io(M, sm, w, a.st, env, true);
}

template <class C>
__global__ void __launch_bounds__(512) k_reset3(const DevModel* __restrict__ G, WResetArgs a) {
  const WModel<C>* Mp; WLane w; int env;
  float* sm = w_setup<C>(G, Mp, w, env, a.n);
  const WModel<C>& M = *Mp;
  if (w.live && a.mask && !a.mask[env]) w.live = false;
  if (!__any_sync(W_FULL, w.live)) return;
  size_t eo = w.live ? (size_t)env : 0;
  const SmplsimEnvCfg& c = M.cfg;
  int init = a.init_mode < 0 ? c.state_init : a.init_mode;
  w_task_io(M, sm, w, a.st, env, false);
  if (w.live && w.li == 0) {
    int* ti = (int*)(sm + C::tsk);
    ((int*)sm)[C::lim] = 0;
    for (int cs = 0; cs < M.nslot; cs++) ((int*)sm)[C::cflag + cs] = 0;
    if (c.task == SMPLSIM_TASK_GETUP) ti[W_TSK_RECOV] = c.recovery_steps;
    if (!c.legacy_change_step) ti[W_TSK_CURT] = 0;
    w_reset_task(M, sm, env);   // sees the old cur_t when legacy_change_step (quirk Q4)
  }
  if (w.live) {
    for (int i = w.li; i < M.nv + 1; i += C::LPE) sm[C::qpos + i] = 0.f;
    for (int i = w.li; i < M.nv; i += C::LPE) { sm[C::qvel + i] = 0.f; sm[C::qacc + i] = 0.f; }
    for (int i = w.li; i < M.nu; i += C::LPE) { sm[C::tau + i] = 0.f; sm[C::act + i] = 0.f; }
  }
  __syncwarp();
  WFwd fo; fo.mask = 0ull; fo.iters = 0; fo.status = 0;
  if (init == SMPLSIM_INIT_MOCAP) {
    w_copy<C>(sm + C::qpos, a.qpos0 + eo * (M.nv + 1), M.nv + 1, w);
    w_copy<C>(sm + C::qvel, a.qvel0 + eo * M.nv, M.nv, w);
  } else if (w.live && w.li == 0) {
    float* q = sm + C::qpos;
    if (init == SMPLSIM_INIT_DEFAULT) { q[2] = 0.94f; q[3] = q[4] = q[5] = q[6] = 0.5f; }
    else { q[2] = 0.3f; q[3] = 1.0f; }
  }
  __syncwarp();
  if (init == SMPLSIM_INIT_FALL) {
    const bool spd_st = (c.control_mode == SMPLSIM_CTRL_UHC_PD && c.spd_stale);
    if (spd_st) w_fk(M, sm, w, true);   // mj_forward: inertia / bias of the initial state
    int ngrp = (M.nu + 3) / 4;
    for (int k3 = 0; k3 < 3; k3++) {
      int* ti = (int*)(sm + C::tsk);
      uint32_t base = w.live ? (uint32_t)ti[W_TSK_RNG] : 0u;
      if (w.live) {
        for (int gidx = w.li; gidx < ngrp; gidx += C::LPE) {
          uint32_t r[4];
          philox4x32(base + (uint32_t)gidx, (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
          for (int j = 0; j < 4 && 4 * gidx + j < M.nu; j++) sm[C::act + 4 * gidx + j] = u01(r[j]) - 0.5f;
        }
      }
      __syncwarp();
      if (w.live && w.li == 0) ti[W_TSK_RNG] = (int)(base + (uint32_t)ngrp);
      __syncwarp();
      if (spd_st) w_spd_prepare(M, sm, w);   // factors of the last forward pass, PD error of the current state and the new action
      w_substeps(M, sm, w, c.nsubsteps, 0, &fo, a.st, env, false, false, 0);
    }
  }
  // reset_sim(): mj_forward at the reset state
  w_fk(M, sm, w, true);
  {
    int nrows = 0;
    unsigned long long m = w_collide(M, sm, w, &nrows);
    unsigned lo = w_gor<C>((unsigned)(m & 0xffffffffull)), hi = w_gor<C>((unsigned)(m >> 32));
    fo.mask = ((unsigned long long)hi << 32) | lo;
  }
  if (w.live) {
    for (int b = w.li; b < M.nb; b += C::LPE) {
      S6 v = ld6(sm + C::vel + 6 * b);
      st3(sm + C::acc2 + 6 * b, v.l + cross(v.a, ld3(sm + C::xpos + 3 * b)));
      st3(sm + C::acc2 + 6 * b + 3, v.a);
    }
    if (w.li == 0) ((int*)(sm + C::tsk))[W_TSK_CURT] = 0;
  }
  __syncwarp();
  w_write_obs(M, sm, w, a.obs ? a.obs + eo * M.obs_dim : nullptr);
  w_write_aux(M, sm, w, a.aux, env, fo);
  w_copy<C>(a.st.qpos + eo * (M.nv + 1), sm + C::qpos, M.nv + 1, w);
  w_copy<C>(a.st.qvel + eo * M.nv, sm + C::qvel, M.nv, w);
  w_copy<C>(a.st.qpos_fwd + eo * (M.nv + 1), sm + C::qpos, M.nv + 1, w);
  w_copy<C>(a.st.qvel_fwd + eo * M.nv, sm + C::qvel, M.nv, w);
  w_copy<C>(a.st.qacc_warm + eo * M.nv, sm + C::qacc, M.nv, w);
  w_task_io(M, sm, w, a.st, env, true);
}
