// chain_kernels.cuh -- v2 hot path: chain-lane ABA stepper (4 lanes per env, 8 envs per warp).
//
// Same mathematics as physics.cuh (v1; kept as the generic fallback and as an A/B reference):
// ABA forward dynamics in world axes about the root origin, stable-PD through ABA with
// armature + h*kd, MuJoCo's soft-constraint problem solved by active-set Newton where each
// Newton system is one ABA with contact terms folded into the touched body, exact line search in
// row space.  What changed is the machine mapping (see chain_model.cuh):
//   * lane = kinematic chain, tree sweeps follow a host-computed list schedule (T steps);
//   * all per-body / per-dof intermediates are lane-private local-memory arrays indexed by the
//     step t, which is uniform across the warp  -> every access is one coalesced 128-B line,
//     cached in L1/L2; shared memory only holds the staged model table and junction mailboxes;
//   * compact, non-inlined device functions (v1 was 310 KB of SASS and instruction-fetch bound).
#pragma once
#include "chain_model.cuh"
#include "dev_model.cuh"

#define CH_FULL 0xffffffffu
#define CH_SOLVER_MAXITER 12
#define CH_LS_MAXITER 24
#define CH_MAXLEDGE 6

template <int TMAX>
struct ChainLoc {          // lane-private scratch (local memory), first index = schedule step t
  float q[TMAX][3], qd[TMAX][3], act[TMAX][3], tau[TMAX][3], qacc[TMAX][3], qstar[TMAX][3], spdab[TMAX][3];
  float xpos[TMAX][3], quat[TMAX][4], ax[TMAX][9], vel[TMAX][6], pb[TMAX][6];
  float U[TMAX][18], Dinv[TMAX][3], u[TMAX][3], acc[TMAX][6], acc2[TMAX][6];
  float edge[CH_MAXLEDGE][27];  // local (same-lane, non-adjacent) articulated-inertia edges
  float ct1[TMAX][3];      // contact tangent of the body's geom
  float cpos[TMAX][CH_MAXC][3], cD[TMAX][CH_MAXC], caref[TMAX][CH_MAXC][4], cphi[TMAX][CH_MAXC][4];
  float cr[TMAX][CH_MAXC][4], cdl[TMAX][CH_MAXC][4];
  int cflag[TMAX][CH_MAXC];   // bit0 valid, bits 1..4 working-set membership of the 4 pyramid rows
  float lD[TMAX][3], laref[TMAX][3], lphi[TMAX][3], lr[TMAX][3], ldl[TMAX][3];
  int lflag[TMAX];            // per dof k: bits (3k): lower side violated, (3k+1): upper side, (3k+2): in working set
  float rootq[4];
};

struct ChainCtx {
  const ChainEntry* tab;   // staged table in shared memory [T][CH_LPE]
  float* mb;               // this env's mailbox block in shared memory
  int c;                   // lane within env (chain id)
  int T;
  bool live;               // env exists
};

__device__ __forceinline__ const ChainEntry& ch_entry(const ChainCtx& x, int t) { return x.tab[t * CH_LPE + x.c]; }

// compact sincos (Cody-Waite + minimax polynomials, |x| < ~1e4); keeps the SASS small, ~1 ulp on [-pi/4, pi/4]
__device__ __forceinline__ void ch_sincos(float x, float* s, float* c) {
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

__device__ __forceinline__ float ch_impedance(const ChainConsts& K, float pm) {
  float x = fabsf(pm) / fmaxf(K.solimp[2], 1e-15f);
  if (x >= 1.f) return K.solimp[1];
  if (x <= 0.f) return K.solimp[0];
  float y, pw = K.solimp[4];
  if (pw == 2.0f) y = (x <= K.solimp[3]) ? K.imp_a * x * x : 1.f - K.imp_b * (1.f - x) * (1.f - x);
  else if (pw < 1.0000001f && pw > 0.9999999f) y = x;
  else y = (x <= K.solimp[3]) ? K.imp_a * __powf(x, pw) : 1.f - K.imp_b * __powf(1.f - x, pw);
  return K.solimp[0] + y * (K.solimp[1] - K.solimp[0]);
}

__device__ __forceinline__ S6 ch_dofS(const ChainEntry& e, const float* ax9, const float* xp, int k) {
  V3 a = ld3(ax9 + 3 * k);
  if (e.kind == CH_KIND_ROOTTRANS) return s6(v3(0.f, 0.f, 0.f), a);
  return s6(a, cross(ld3(xp), a));
}

__device__ __forceinline__ void ch_rigid10(const ChainEntry& e, const float* quat, const float* xp, float* r10) {
  Q4 q; q.w = quat[0]; q.x = quat[1]; q.y = quat[2]; q.z = quat[3];
  float R[9];
  q2mat(q, R);
  const float* in = e.inertia;
  float m = e.mass;
  V3 r = ld3(xp) + mrot(R, ld3(e.ipos));
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

__device__ __forceinline__ S6 ch_wrench(const ChainConsts& K, V3 cp, V3 t1, int k) {
  V3 n = ld3(K.plane_n);
  V3 t = (k < 2) ? t1 : cross(n, t1);
  float sg = (k & 1) ? -K.mu : K.mu;
  V3 dir = n + sg * t;
  return s6(cross(cp, dir), dir);
}

// ------------------------------------------------------------------ kinematics / velocities / bias forces (outward sweep)
template <int TMAX>
__device__ __noinline__ void ch_fk(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool vel) {
  for (int t = x.T - 1; t >= 0; t--) {
    const ChainEntry& e = ch_entry(x, t);
    if (x.live && e.pb >= 0) {
      Q4 qc; V3 xp; S6 v, ab;
      v = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)); ab = v;
      if (e.kind == CH_KIND_ROOTTRANS) {
        xp = v3(0.f, 0.f, 0.f);
        qc.w = 1.f; qc.x = qc.y = qc.z = 0.f;
        st3(&L.ax[t][0], v3(1.f, 0.f, 0.f)); st3(&L.ax[t][3], v3(0.f, 1.f, 0.f)); st3(&L.ax[t][6], v3(0.f, 0.f, 1.f));
        if (vel) { v.l = ld3(L.qd[t]); ab.l = v3(-K.grav[0], -K.grav[1], -K.grav[2]); }
      } else {
        Q4 qp;
        if (e.par_t >= 0) {
          xp = ld3(L.xpos[e.par_t]);
          qp.w = L.quat[e.par_t][0]; qp.x = L.quat[e.par_t][1]; qp.y = L.quat[e.par_t][2]; qp.z = L.quat[e.par_t][3];
          if (vel) { v = ld6(L.vel[e.par_t]); ab = ld6(L.acc[e.par_t]); }   // bias acceleration parked in acc during FK
        } else {
          const float* m = x.mb + 19 * e.par_mbox;
          xp = ld3(m);
          qp.w = m[3]; qp.x = m[4]; qp.y = m[5]; qp.z = m[6];
          if (vel) { v = ld6(m + 7); ab = ld6(m + 13); }
        }
        if (e.kind == CH_KIND_ROOTROT) {
          qc.w = L.rootq[0]; qc.x = L.rootq[1]; qc.y = L.rootq[2]; qc.z = L.rootq[3];
          qc = qnormalize(qc);
          L.rootq[0] = qc.w; L.rootq[1] = qc.x; L.rootq[2] = qc.y; L.rootq[3] = qc.z;
          float R[9];
          q2mat(qc, R);
          V3 c0 = v3(R[0], R[3], R[6]), c1 = v3(R[1], R[4], R[7]), c2 = v3(R[2], R[5], R[8]);
          st3(&L.ax[t][0], c0); st3(&L.ax[t][3], c1); st3(&L.ax[t][6], c2);
          if (vel) {
            V3 w = L.qd[t][0] * c0 + L.qd[t][1] * c1 + L.qd[t][2] * c2;
            ab.l = ab.l + cross(v.l, w);      // all three rotational cdof_dot use the translational velocity (mj_comVel)
            v.a = w;
          }
        } else {
          Q4 qpar = qp;
          float Rp[9];
          q2mat(qpar, Rp);
          xp = xp + mrot(Rp, ld3(e.bpos));
          Q4 qb; qb.w = e.bquat[0]; qb.x = e.bquat[1]; qb.y = e.bquat[2]; qb.z = e.bquat[3];
          qc = qmul(qpar, qb);
          for (int k = 0; k < e.ndof; k++) {
            V3 al = ld3(e.axis + 3 * k);
            V3 a = qrot(qc, al);
            st3(&L.ax[t][3 * k], a);
            if (vel) {
              S6 S = s6(a, cross(xp, a));
              float qd = L.qd[t][k];
              S6 sd = cross_motion(v, S);
              ab = ab + qd * sd;
              v = v + qd * S;
            }
            float sn, cs;
            ch_sincos(0.5f * L.q[t][k], &sn, &cs);
            Q4 qj; qj.w = cs; qj.x = al.x * sn; qj.y = al.y * sn; qj.z = al.z * sn;
            qc = qmul(qc, qj);
          }
          qc = qnormalize(qc);
        }
      }
      st3(L.xpos[t], xp);
      L.quat[t][0] = qc.w; L.quat[t][1] = qc.x; L.quat[t][2] = qc.y; L.quat[t][3] = qc.z;
      if (vel) {
        st6(L.vel[t], v); st6(L.acc[t], ab);
        S6 f = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
        if (e.kind != CH_KIND_ROOTTRANS) {
          float r10[10];
          ch_rigid10(e, L.quat[t], L.xpos[t], r10);
          f = rb_mul(r10, ab) + cross_force(v, rb_mul(r10, v));
        }
        st6(L.pb[t], f);
      }
      if (e.out_mbox >= 0) {
        float* m = x.mb + 19 * e.out_mbox;
        st3(m, xp); m[3] = qc.w; m[4] = qc.x; m[5] = qc.y; m[6] = qc.z;
        if (vel) { st6(m + 7, v); st6(m + 13, ab); }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ABA inward sweep
// flags: 1 INERTIA (build U, 1/D) | 2 FORCE (u, pA) | 4 PB (bias forces in) | 8 CONTACTS (working-set rows)
// tmode (FORCE): 0 joint force = tau (+ limit rows), 1 zero, 2 stable-PD  -kp e - kd qd
// dmode (INERTIA): 0 armature (+ limit rows), 1 armature + h kd
#define CH_INERTIA 1
#define CH_FORCE 2
#define CH_PB 4
#define CH_CONTACTS 8
template <int TMAX>
__device__ __noinline__ void ch_inward(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool run, int flags, int tmode, int dmode) {
  float A[21];
  S6 p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
#pragma unroll
  for (int j = 0; j < 21; j++) A[j] = 0.f;
  const bool inertia = flags & CH_INERTIA, force = flags & CH_FORCE;
  float* mbe = x.mb + 19 * K.n_mbox;   // cross-lane edge mailboxes (27 words each) follow the FK mailboxes
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (run && e.pb >= 0) {
      if (!e.carry_in) {
#pragma unroll
        for (int j = 0; j < 21; j++) A[j] = 0.f;
        p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
      }
      if (e.kind != CH_KIND_ROOTTRANS) {
        if (inertia) {
          float r10[10], B[21];
          ch_rigid10(e, L.quat[t], L.xpos[t], r10);
          rb_expand(r10, B);
#pragma unroll
          for (int j = 0; j < 21; j++) A[j] += B[j];
        }
        if (force && (flags & CH_PB)) p = p + ld6(L.pb[t]);
        if ((flags & CH_CONTACTS) && e.geom >= 0) {
          V3 t1 = ld3(L.ct1[t]);
          for (int s = 0; s < CH_MAXC; s++) {
            int fl = L.cflag[t][s];
            if (!(fl & 1) || !(fl & 30)) continue;
            V3 cp = ld3(L.cpos[t][s]);
            float D = L.cD[t][s];
            for (int k = 0; k < 4; k++) {
              if (!(fl & (2 << k))) continue;
              S6 xw = ch_wrench(K, cp, t1, k);
              float xv[6] = {xw.a.x, xw.a.y, xw.a.z, xw.l.x, xw.l.y, xw.l.z};
              if (inertia) sym_rank1(A, xv, -D);
              if (force) p = p - (D * L.caref[t][s][k]) * xw;
            }
          }
        }
      }
      for (int j = 0; j < 3; j++) {
        int ed = e.in_edge[j];
        if (ed < 0) continue;
        const float* src = (ed >= CH_EDGE_MBOX) ? mbe + 27 * (ed - CH_EDGE_MBOX) : L.edge[ed];
        if (inertia) {
#pragma unroll
          for (int i = 0; i < 21; i++) A[i] += src[i];
        }
        if (force) p = p + ld6(src + 21);
      }
      for (int k = e.ndof - 1; k >= 0; k--) {
        S6 S = ch_dofS(e, L.ax[t], L.xpos[t], k);
        float s[6] = {S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z}, Uv[6], di;
        int lf = (L.lflag[t] >> (3 * k)) & 7;
        if (inertia) {
          sym_mul(A, s, Uv);
          float D = e.arm[k];
          if (dmode == 1) D += K.h * e.kd[k];
          else if (lf & 4) D += L.lD[t][k];
#pragma unroll
          for (int j = 0; j < 6; j++) D = fmaf(s[j], Uv[j], D);
          di = 1.0f / D;
#pragma unroll
          for (int j = 0; j < 6; j++) L.U[t][6 * k + j] = Uv[j];
          L.Dinv[t][k] = di;
          sym_rank1(A, Uv, di);
        } else {
#pragma unroll
          for (int j = 0; j < 6; j++) Uv[j] = L.U[t][6 * k + j];
          di = L.Dinv[t][k];
        }
        if (force) {
          float tin = 0.f;
          if (tmode == 0) {
            tin = L.tau[t][k];
            if (lf & 4) tin += ((lf & 1) ? 1.f : -1.f) * L.lD[t][k] * L.laref[t][k];
          } else if (tmode == 2 && e.kind == CH_KIND_HINGE) {
            float tgt = fmaf(L.act[t][k], e.ascale[k], e.aoffset[k]);
            float err = L.q[t][k] + L.qd[t][k] * K.h - tgt;
            tin = -e.kp[k] * err - e.kd[k] * L.qd[t][k];
          }
          float uu = tin - dot6(S, p);
          L.u[t][k] = uu;
          p = p + (uu * di) * s6(v3(Uv[0], Uv[1], Uv[2]), v3(Uv[3], Uv[4], Uv[5]));
        }
      }
      if (!e.carry_out && e.out_edge >= 0) {
        float* dst = (e.out_edge >= CH_EDGE_MBOX) ? mbe + 27 * (e.out_edge - CH_EDGE_MBOX) : L.edge[e.out_edge];
        if (inertia) {
#pragma unroll
          for (int i = 0; i < 21; i++) dst[i] = A[i];
        }
        if (force) st6(dst + 21, p);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ABA outward sweep: qdd = (u - U.a)/D ; a += S qdd
// which: 0 -> qacc, 1 -> qstar, 2 -> spdab ; accsel: 0 -> acc, 1 -> acc2 ; mode 1: accumulate S*q only (q from qacc), no solve
template <int TMAX>
__device__ __noinline__ void ch_outward(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool run, int which, int mode) {
  for (int t = x.T - 1; t >= 0; t--) {
    const ChainEntry& e = ch_entry(x, t);
    if (run && e.pb >= 0) {
      S6 a = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
      if (e.kind != CH_KIND_ROOTTRANS) a = (e.par_t >= 0) ? ld6(L.acc[e.par_t]) : ld6(x.mb + 19 * e.par_mbox);
      for (int k = 0; k < e.ndof; k++) {
        S6 S = ch_dofS(e, L.ax[t], L.xpos[t], k);
        float qdd;
        if (mode == 1) qdd = L.qacc[t][k];
        else {
          qdd = L.Dinv[t][k] * (L.u[t][k] - dot6(ld6(&L.U[t][6 * k]), a));
          if (which == 0) L.qacc[t][k] = qdd; else if (which == 1) L.qstar[t][k] = qdd; else L.spdab[t][k] = qdd;
        }
        a = a + qdd * S;
      }
      st6(L.acc[t], a);
      if (e.out_mbox >= 0) st6(x.mb + 19 * e.out_mbox, a);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ collision + limit rows at the current state
template <int TMAX>
__device__ __noinline__ unsigned long long ch_collide(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, float root_h, int* nrows_out) {
  unsigned long long mask = 0ull;
  int nrows = 0;
  V3 n = ld3(K.plane_n);
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (!x.live || e.pb < 0) continue;
    // ---- joint limits (hinges, margin 0)
    int lfl = 0;
    if (e.kind == CH_KIND_HINGE && e.limited) {
      for (int k = 0; k < e.ndof; k++) {
        if (!((e.limited >> k) & 1)) continue;
        float q = L.q[t][k], dlo = q - e.range[2 * k], dhi = e.range[2 * k + 1] - q, dist = 0.f, sg = 0.f;
        int f = 0;
        if (dlo < 0.f) { dist = dlo; sg = 1.f; f = 1; }
        else if (dhi < 0.f) { dist = dhi; sg = -1.f; f = 2; }
        if (f) {
          float imp = ch_impedance(K, dist);
          L.lD[t][k] = 1.0f / fmaxf((1.f - imp) / imp * e.diw0[k], 1e-15f);
          L.laref[t][k] = -K.B * sg * L.qd[t][k] - K.K * imp * dist;
          lfl |= f << (3 * k);
          nrows++;
        }
      }
    }
    L.lflag[t] = lfl;
    // ---- floor plane vs this body's geom
#pragma unroll
    for (int s = 0; s < CH_MAXC; s++) L.cflag[t][s] = 0;
    if (e.geom < 0) continue;
    Q4 q; q.w = L.quat[t][0]; q.x = L.quat[t][1]; q.y = L.quat[t][2]; q.z = L.quat[t][3];
    float R[9];
    q2mat(q, R);
    V3 xb = ld3(L.xpos[t]);
    V3 c = xb + mrot(R, ld3(e.gpos));
    float d0 = root_h + dot(n, c);
    int cnt = 0;
    V3 t1 = ld3(K.t1_default);
    float dist_s[CH_MAXC];
    V3 cp_s[CH_MAXC];
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
      st3(L.ct1[t], t1);
      S6 v = ld6(L.vel[t]);
      for (int s = 0; s < cnt; s++) {
        st3(L.cpos[t][s], cp_s[s]);
        float pm = dist_s[s] - K.margin, imp = ch_impedance(K, pm);
        float R0 = fmaxf((1.f - imp) / imp * (e.tran_iw0 + K.mu * K.mu * e.tran_iw0), 1e-15f);
        float R1 = R0 / fmaxf(K.impratio, 1e-15f), mu = K.mu * sqrtf(R1 / R0);
        L.cD[t][s] = 1.0f / (2.f * mu * mu * R0);
        float kterm = K.K * imp * pm;
        for (int k = 0; k < 4; k++) L.caref[t][s][k] = -K.B * dot6(ch_wrench(K, cp_s[s], t1, k), v) - kterm;
        L.cflag[t][s] = 1;
      }
      mask |= 1ull << (e.geom + 1);
      nrows += 4 * cnt;
    }
  }
  *nrows_out = nrows;
  return mask;
}

// sum over the 4 lanes of an env
__device__ __forceinline__ float ch_envsum(float v) {
  v += __shfl_xor_sync(CH_FULL, v, 1);
  v += __shfl_xor_sync(CH_FULL, v, 2);
  return v;
}
__device__ __forceinline__ bool ch_envall(bool p, int lane) {
  unsigned m = __ballot_sync(CH_FULL, p);
  return ((m >> (lane & ~3)) & 0xFu) == 0xFu;
}
__device__ __forceinline__ bool ch_envany(bool p, int lane) {
  unsigned m = __ballot_sync(CH_FULL, p);
  return ((m >> (lane & ~3)) & 0xFu) != 0u;
}

// ------------------------------------------------------------------ rows at the acceleration field L.acc / dof accelerations qd
// mode 0: working set := (r < 0) with qd = qacc (warm start);  mode 1: rs from qstar, store cdl = rs (temporarily), report sameness
template <int TMAX>
__device__ __noinline__ bool ch_eval_rows(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool run, int mode, int lane) {
  bool same = true;
  if (run) {
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb < 0) continue;
      if (e.geom >= 0) {
        S6 a = ld6(L.acc[t]);
        V3 t1 = ld3(L.ct1[t]);
        for (int s = 0; s < CH_MAXC; s++) {
          int fl = L.cflag[t][s];
          if (!(fl & 1)) continue;
          V3 cp = ld3(L.cpos[t][s]);
          int nf = 1;
          for (int k = 0; k < 4; k++) {
            float r = dot6(ch_wrench(K, cp, t1, k), a) - L.caref[t][s][k];
            if (mode == 1) L.cdl[t][s][k] = r;
            if (r < 0.f) nf |= 2 << k;
          }
          if (mode == 0) L.cflag[t][s] = nf;
          else if (nf != fl) same = false;
        }
      }
      int lf = L.lflag[t];
      if (lf) {
        int nlf = lf;
        for (int k = 0; k < e.ndof; k++) {
          int f = (lf >> (3 * k)) & 7;
          if (!(f & 3)) continue;
          float sg = (f & 1) ? 1.f : -1.f;
          float r = sg * (mode == 0 ? L.qacc[t][k] : L.qstar[t][k]) - L.laref[t][k];
          if (mode == 1) L.ldl[t][k] = r;
          int nf = (f & 3) | ((r < 0.f) ? 4 : 0);
          nlf = (nlf & ~(7 << (3 * k))) | (nf << (3 * k));
        }
        if (mode == 0) L.lflag[t] = nlf;
        else if (nlf != lf) same = false;
      }
    }
  }
  return ch_envall(same, lane);
}

// after eval_rows(mode 1): cdl/ldl hold rs.  adopt==true: first iterate (r := rs, phi := -D rs on the used set, set := rs<0).
// adopt==false: cdl := rs - r (direction), and g1, g2 of the line search are accumulated.
template <int TMAX>
__device__ __noinline__ void ch_rows_update(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool run, bool adopt, float* g1o, float* g2o) {
  float g1 = 0.f, g2 = 0.f;
  if (run) {
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb < 0) continue;
      if (e.geom >= 0) {
        for (int s = 0; s < CH_MAXC; s++) {
          int fl = L.cflag[t][s];
          if (!(fl & 1)) continue;
          float D = L.cD[t][s];
          int nf = 1;
          for (int k = 0; k < 4; k++) {
            float rs = L.cdl[t][s][k];
            float phs = (fl & (2 << k)) ? -D * rs : 0.f;
            if (adopt) {
              L.cr[t][s][k] = rs; L.cphi[t][s][k] = phs;
              if (rs < 0.f) nf |= 2 << k;
            } else {
              float d = rs - L.cr[t][s][k], ph = L.cphi[t][s][k];
              L.cdl[t][s][k] = d;
              g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
            }
          }
          if (adopt) L.cflag[t][s] = nf;
        }
      }
      int lf = L.lflag[t];
      if (lf) {
        int nlf = lf;
        for (int k = 0; k < e.ndof; k++) {
          int f = (lf >> (3 * k)) & 7;
          if (!(f & 3)) continue;
          float rs = L.ldl[t][k], phs = (f & 4) ? -L.lD[t][k] * rs : 0.f;
          if (adopt) {
            L.lr[t][k] = rs; L.lphi[t][k] = phs;
            int nf = (f & 3) | ((rs < 0.f) ? 4 : 0);
            nlf = (nlf & ~(7 << (3 * k))) | (nf << (3 * k));
          } else {
            float d = rs - L.lr[t][k], ph = L.lphi[t][k];
            L.ldl[t][k] = d;
            g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
          }
        }
        if (adopt) L.lflag[t] = nlf;
      }
    }
  }
  *g1o = g1; *g2o = g2;
}

// line-search partial sums at step length al (this lane's rows); al < 0: apply the step -al instead (update r, phi, working set)
template <int TMAX>
__device__ __noinline__ void ch_ls(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool run, float al, float* s1o, float* s2o) {
  float s1 = 0.f, s2 = 0.f;
  const bool apply = al < 0.f;
  if (apply) al = -al;
  if (run) {
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb < 0) continue;
      if (e.geom >= 0) {
        for (int s = 0; s < CH_MAXC; s++) {
          int fl = L.cflag[t][s];
          if (!(fl & 1)) continue;
          float D = L.cD[t][s];
          int nf = 1;
          for (int k = 0; k < 4; k++) {
            float r = L.cr[t][s][k], d = L.cdl[t][s][k], v = fmaf(al, d, r);
            if (apply) {
              float rs = r + d, ph = L.cphi[t][s][k], phs = (fl & (2 << k)) ? -D * rs : 0.f;
              L.cr[t][s][k] = v; L.cphi[t][s][k] = fmaf(al, phs - ph, ph);
              if (v < 0.f) nf |= 2 << k;
            } else if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
          }
          if (apply) L.cflag[t][s] = nf;
        }
      }
      int lf = L.lflag[t];
      if (lf) {
        int nlf = lf;
        for (int k = 0; k < e.ndof; k++) {
          int f = (lf >> (3 * k)) & 7;
          if (!(f & 3)) continue;
          float D = L.lD[t][k], r = L.lr[t][k], d = L.ldl[t][k], v = fmaf(al, d, r);
          if (apply) {
            float rs = r + d, ph = L.lphi[t][k], phs = (f & 4) ? -D * rs : 0.f;
            L.lr[t][k] = v; L.lphi[t][k] = fmaf(al, phs - ph, ph);
            int nf = (f & 3) | ((v < 0.f) ? 4 : 0);
            nlf = (nlf & ~(7 << (3 * k))) | (nf << (3 * k));
          } else if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
        }
        if (apply) L.lflag[t] = nlf;
      }
    }
  }
  *s1o = s1; *s2o = s2;
}

// ------------------------------------------------------------------ constraint solve; returns #extra ABA solves of this env
template <int TMAX>
__device__ __noinline__ int ch_solve(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, bool any_rows, int lane) {
  // envs without rows: one plain ABA
  bool plain = x.live && !any_rows;
  if (__any_sync(CH_FULL, plain)) {
    ch_inward(K, x, L, plain, CH_INERTIA | CH_FORCE | CH_PB, 0, 0);
    ch_outward(K, x, L, plain, 0, 0);
  }
  bool run = x.live && any_rows;
  if (!__any_sync(CH_FULL, run)) return 0;
  ch_outward(K, x, L, run, 0, 1);                 // body accelerations of the warm start
  ch_eval_rows(K, x, L, run, 0, lane);
  bool have_point = false;
  int it = 0, iters = 0;
  for (; it < CH_SOLVER_MAXITER; it++) {
    if (!__any_sync(CH_FULL, run)) break;
    ch_inward(K, x, L, run, CH_INERTIA | CH_FORCE | CH_PB | CH_CONTACTS, 0, 0);
    ch_outward(K, x, L, run, 1, 0);
    bool same = ch_eval_rows(K, x, L, run, 1, lane);
    bool fin = run && same;
    bool adopt = run && !same && !have_point;
    bool lsrch = run && !same && have_point;
    if (fin || adopt) {
      for (int t = 0; t < x.T; t++)
#pragma unroll
        for (int k = 0; k < 3; k++) L.qacc[t][k] = L.qstar[t][k];
    }
    if (adopt) {
      for (int t = 0; t < x.T; t++) st6(L.acc2[t], ld6(L.acc[t]));
    }
    float g1, g2;
    ch_rows_update(K, x, L, adopt || lsrch, adopt, &g1, &g2);
    if (fin) { run = false; iters = it; }
    if (adopt) have_point = true;
    if (__any_sync(CH_FULL, lsrch)) {
      g1 = ch_envsum(g1); g2 = ch_envsum(g2);
      float s1, s2;
      ch_ls(K, x, L, lsrch, 0.f, &s1, &s2);
      s1 = ch_envsum(s1);
      float f0 = g1 + s1, al = 0.f, lo = 0.f, hi = -1.f, tol = 1e-6f * fabsf(f0);
      bool searching = lsrch && (f0 < -1e-4f * (fabsf(g1) + fabsf(s1)));
      if (searching) al = 1.f;
      for (int ls = 0; ls < CH_LS_MAXITER; ls++) {
        if (!__any_sync(CH_FULL, searching)) break;
        ch_ls(K, x, L, searching, al, &s1, &s2);
        s1 = ch_envsum(s1); s2 = ch_envsum(s2);
        if (searching) {
          float f = g1 + al * g2 + s1, fp = g2 + s2;
          if (fabsf(f) <= tol) searching = false;
          else {
            if (f < 0.f) lo = al; else hi = al;
            float an = (fp > 0.f) ? al - f / fp : -1.f;
            if (!(an > lo) || (hi > 0.f && !(an < hi))) an = (hi > 0.f) ? 0.5f * (lo + hi) : 2.f * al;
            an = fminf(an, 16.f);
            if (an == al) searching = false; else al = an;
          }
        }
      }
      bool step = lsrch && (al > 0.f);
      if (lsrch && !step) { run = false; iters = it; }   // no descent left: keep the current iterate
      if (step) {
        for (int t = 0; t < x.T; t++) {
#pragma unroll
          for (int k = 0; k < 3; k++) L.qacc[t][k] = fmaf(al, L.qstar[t][k] - L.qacc[t][k], L.qacc[t][k]);
          S6 a2 = ld6(L.acc2[t]);
          st6(L.acc2[t], a2 + al * (ld6(L.acc[t]) - a2));
        }
      }
      ch_ls(K, x, L, step, -al, &s1, &s2);
    }
  }
  if (run) iters = it;
  return iters;
}

// ------------------------------------------------------------------ stable PD
template <int TMAX>
__device__ __noinline__ void ch_spd_prepare(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L) {
  ch_inward(K, x, L, x.live, CH_INERTIA | CH_FORCE | CH_PB, 1, 1);   // (M + h Kd) factors; a_bias = (M + h Kd)^-1 (-C)
  ch_outward(K, x, L, x.live, 2, 0);
}

template <int TMAX>
__device__ __noinline__ void ch_torque(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L) {
  int mode = K.cfg.control_mode;
  if (mode == SMPLSIM_CTRL_UHC_PD) {
    ch_inward(K, x, L, x.live, CH_FORCE, 2, 0);
    ch_outward(K, x, L, x.live, 1, 0);
  }
  if (!x.live) return;
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0 || e.kind != CH_KIND_HINGE) continue;
    for (int k = 0; k < e.ndof; k++) {
      float a = L.act[t][k], tq;
      if (mode == SMPLSIM_CTRL_TORQUE) tq = a * e.ascale[k];
      else {
        float tgt = fmaf(a, e.ascale[k], e.aoffset[k]);
        if (mode == SMPLSIM_CTRL_PD) tq = -e.kp[k] * (L.q[t][k] - tgt) - e.kd[k] * L.qd[t][k];
        else {
          float err = L.q[t][k] + L.qd[t][k] * K.h - tgt;
          tq = -e.kp[k] * err - e.kd[k] * (L.qd[t][k] + (L.spdab[t][k] + L.qstar[t][k]) * K.h);
        }
      }
      L.tau[t][k] = fminf(fmaxf(tq, -e.tlim[k]), e.tlim[k]);
    }
  }
}

// ------------------------------------------------------------------ semi-implicit Euler; returns this lane's root displacement (root-trans lane)
template <int TMAX>
__device__ __noinline__ V3 ch_integrate(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, float* rootpos) {
  V3 disp = v3(0.f, 0.f, 0.f);
  if (!x.live) return disp;
  float h = K.h;
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0) continue;
    for (int k = 0; k < e.ndof; k++) L.qd[t][k] = fmaf(h, L.qacc[t][k], L.qd[t][k]);
    if (e.kind == CH_KIND_HINGE) {
      for (int k = 0; k < e.ndof; k++) L.q[t][k] = fmaf(h, L.qd[t][k], L.q[t][k]);
    } else if (e.kind == CH_KIND_ROOTTRANS) {
      disp = h * ld3(L.qd[t]);
      st3(L.q[t], ld3(L.q[t]) + disp);
      st3(rootpos, ld3(L.q[t]));
    } else {
      V3 w = ld3(L.qd[t]);
      float n = sqrtf(dot(w, w)), ang = n * h;
      Q4 q; q.w = L.rootq[0]; q.x = L.rootq[1]; q.y = L.rootq[2]; q.z = L.rootq[3];
      if (ang > 0.f) {
        float sn, cs;
        ch_sincos(0.5f * ang, &sn, &cs);
        float s = sn / n;
        Q4 dq; dq.w = cs; dq.x = w.x * s; dq.y = w.y * s; dq.z = w.z * s;
        q = qmul(q, dq);
      }
      q = qnormalize(q);
      L.rootq[0] = q.w; L.rootq[1] = q.x; L.rootq[2] = q.y; L.rootq[3] = q.z;
    }
  }
  return disp;
}

// ====================================================================================================================
// env-level kernels (v2)
// ====================================================================================================================
#define CH_SC_ROOTPOS 0   // per-env scalars in shared memory (after the mailboxes)
#define CH_SC_HQ 3
#define CH_SC_TSK 8       // target[4], change_step, cur_t, recovery, rng
#define CH_SC_WORDS 16

struct ChainStepArgs {
  SmplsimState st;
  SmplsimAux aux;
  const float* action;
  float* obs;
  float* reward;
  uint8_t* terminated;
  uint8_t* truncated;
  int n, nsub, mode;
};
struct ChainResetArgs {
  SmplsimState st;
  SmplsimAux aux;
  const uint8_t* mask;
  const float* qpos0;
  const float* qvel0;
  float* obs;
  int n, init_mode;
};

// what: bit0 qpos/qvel from (qp, qv) ; bit1 warm start -> qacc ; bit2 action -> act ; bit3 action -> tau
template <int TMAX>
__device__ __noinline__ void ch_load(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, const float* qp, const float* qv, const float* qw,
                                     const float* a, int what, float* sc) {
  if (!x.live) return;
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0) continue;
    for (int k = 0; k < e.ndof; k++) {
      int d = e.dofadr + k;
      if (what & 1) {
        L.qd[t][k] = qv[d];
        if (e.kind == CH_KIND_HINGE) L.q[t][k] = qp[d + 1];
        else if (e.kind == CH_KIND_ROOTTRANS) L.q[t][k] = qp[k];
      }
      if (what & 2) L.qacc[t][k] = qw[d];
      if (e.kind == CH_KIND_HINGE) {
        if (what & 4) L.act[t][k] = a[d - 6];
        if (what & 8) L.tau[t][k] = a[d - 6];
      }
    }
    if (what & 1) {
      if (e.kind == CH_KIND_ROOTROT) { L.rootq[0] = qp[3]; L.rootq[1] = qp[4]; L.rootq[2] = qp[5]; L.rootq[3] = qp[6]; }
      if (e.kind == CH_KIND_ROOTTRANS) st3(sc + CH_SC_ROOTPOS, ld3(qp));
    }
  }
}

template <int TMAX>
__device__ __noinline__ void ch_store(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, float* qp, float* qv, float* qw) {
  if (!x.live) return;
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0) continue;
    for (int k = 0; k < e.ndof; k++) {
      int d = e.dofadr + k;
      qv[d] = L.qd[t][k];
      if (qw) qw[d] = L.qacc[t][k];
      if (e.kind == CH_KIND_HINGE) qp[d + 1] = L.q[t][k];
      else if (e.kind == CH_KIND_ROOTTRANS) qp[k] = L.q[t][k];
    }
    if (e.kind == CH_KIND_ROOTROT) { qp[3] = L.rootq[0]; qp[4] = L.rootq[1]; qp[5] = L.rootq[2]; qp[6] = L.rootq[3]; }
  }
}

// reset_task (lane c == 0 of the env)
__device__ __forceinline__ void ch_reset_task(const ChainConsts& K, float* sc, int env) {
  const SmplsimEnvCfg& c = K.cfg;
  if (c.task == SMPLSIM_TASK_NONE) return;
  float* t = sc + CH_SC_TSK;
  int* ti = (int*)t;
  uint32_t r[4];
  philox4x32((uint32_t)ti[7], (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
  ti[7] = ti[7] + 1;
  if (c.task == SMPLSIM_TASK_SPEED) t[0] = (float)(c.tar_speed_max - c.tar_speed_min) * u01(r[0]) + (float)c.tar_speed_min;
  else if (c.task == SMPLSIM_TASK_REACH) {
    t[0] = (float)c.tar_dist_max * (2.0f * u01(r[0]) - 1.0f);
    t[1] = (float)c.tar_dist_max * (2.0f * u01(r[1]) - 1.0f);
    t[2] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[2]) + (float)c.tar_height_min;
  } else t[0] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[0]) + (float)c.tar_height_min;
  ti[4] = ti[5] + rand_range(r[3], c.change_steps_min, c.change_steps_max);
}
__device__ __forceinline__ void ch_task_io(const ChainCtx& x, float* sc, const SmplsimState& st, int env, bool store) {
  if (!x.live || x.c != 0) return;
  float* t = sc + CH_SC_TSK;
  int* ti = (int*)t;
  if (!store) {
    for (int j = 0; j < 4; j++) t[j] = st.task_target[4 * env + j];
    ti[4] = st.task_change_step[env]; ti[5] = st.progress[env]; ti[6] = st.recovery[env]; ti[7] = (int)st.rng_counter[env];
  } else {
    for (int j = 0; j < 4; j++) st.task_target[4 * env + j] = t[j];
    st.task_change_step[env] = ti[4]; st.progress[env] = ti[5]; st.recovery[env] = ti[6]; st.rng_counter[env] = (uint32_t)ti[7];
  }
}

struct ChainFwd { unsigned long long mask; int iters; };

// nsub x [torque + mj_step] on the staged state
template <int TMAX>
__device__ __noinline__ V3 ch_substeps(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, float* sc, int lane, int nsub, int raw, ChainFwd* fo,
                                       const SmplsimState& st, int env, bool write_fwd, bool prep_last) {
  const bool spd = (K.cfg.control_mode == SMPLSIM_CTRL_UHC_PD), stale = K.cfg.spd_stale != 0;
  V3 disp = v3(0.f, 0.f, 0.f);
  for (int s = 0; s < nsub; s++) {
    bool did_fk = false;
    if (!raw) {
      if (spd && !stale) { ch_fk(K, x, L, true); ch_spd_prepare(K, x, L); did_fk = true; }
      ch_torque(K, x, L);
    }
    if (!did_fk) ch_fk(K, x, L, true);
    float root_h = dot(ld3(K.plane_n), ld3(sc + CH_SC_ROOTPOS) - ld3(K.plane_pos));
    int nrows = 0;
    unsigned long long m = ch_collide(K, x, L, root_h, &nrows);
    unsigned lo = (unsigned)(m & 0xffffffffull), hi = (unsigned)(m >> 32);
    lo |= __shfl_xor_sync(CH_FULL, lo, 1); lo |= __shfl_xor_sync(CH_FULL, lo, 2);
    hi |= __shfl_xor_sync(CH_FULL, hi, 1); hi |= __shfl_xor_sync(CH_FULL, hi, 2);
    fo->mask = ((unsigned long long)hi << 32) | lo;
    bool any_rows = ch_envany(nrows > 0, lane);
    fo->iters = ch_solve(K, x, L, any_rows, lane);
    if (s == nsub - 1) {
      if (x.live) {
        for (int t = 0; t < x.T; t++) {     // framelinvel / frameangvel of the last forward pass (quirk Q2) parked in acc2
          const ChainEntry& e = ch_entry(x, t);
          if (e.pb < 0) continue;
          S6 v = ld6(L.vel[t]);
          st3(&L.acc2[t][0], v.l + cross(v.a, ld3(L.xpos[t])));
          st3(&L.acc2[t][3], v.a);
        }
        if (write_fwd) ch_store(K, x, L, st.qpos_fwd + (size_t)env * K.nq, st.qvel_fwd + (size_t)env * K.nv, nullptr);
      }
    }
    if (spd && stale && !raw && (s < nsub - 1 || prep_last)) ch_spd_prepare(K, x, L);
    V3 d = ch_integrate(K, x, L, sc + CH_SC_ROOTPOS);
    disp = disp + d;
    __syncwarp();
  }
  return disp;
}

__device__ __forceinline__ Q4 ch_heading_inv(const ChainConsts& K, Q4 root) {
  if (!K.cfg.upright_start) { Q4 bc; bc.w = 0.5f; bc.x = -0.5f; bc.y = -0.5f; bc.z = -0.5f; root = qmul(root, bc); }
  V3 rd = qrot_ref(root, v3(1.f, 0.f, 0.f));
  float hd = atan2f(rd.y, rd.x), sn, cs;
  ch_sincos(-0.5f * hd, &sn, &cs);
  Q4 h; h.w = cs; h.x = 0.f; h.y = 0.f; h.z = sn;
  return qnormalize(h);
}

// compute_observations: each lane writes the entries of its own bodies straight to the obs row
template <int TMAX>
__device__ __noinline__ void ch_write_obs(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, float* sc, float* ob) {
  // heading from the root-rot lane
  if (x.live) {
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb >= 0 && e.kind == CH_KIND_ROOTROT) {
        Q4 r0; r0.w = L.rootq[0]; r0.x = L.rootq[1]; r0.y = L.rootq[2]; r0.z = L.rootq[3];
        Q4 h = ch_heading_inv(K, r0);
        sc[CH_SC_HQ] = h.w; sc[CH_SC_HQ + 1] = h.x; sc[CH_SC_HQ + 2] = h.y; sc[CH_SC_HQ + 3] = h.z;
      }
    }
  }
  __syncwarp();
  if (!x.live || !ob) return;
  Q4 hq; hq.w = sc[CH_SC_HQ]; hq.x = sc[CH_SC_HQ + 1]; hq.y = sc[CH_SC_HQ + 2]; hq.z = sc[CH_SC_HQ + 3];
  int nb = K.nb, o = K.cfg.root_height_obs ? 1 : 0, o_rot = o + 3 * (nb - 1), o_vel = o_rot + 6 * nb;
  const float* tk = sc + CH_SC_TSK;
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0) continue;
    if (e.kind == CH_KIND_ROOTTRANS) {
      if (o) ob[0] = sc[CH_SC_ROOTPOS + 2];
      if (K.cfg.self_obs_v == 1) st3(ob + o_vel, qrot_ref(hq, ld3(L.qd[t])));
      continue;
    }
    int b = e.body;
    if (b > 0) st3(ob + o + 3 * (b - 1), qrot_ref(hq, ld3(L.xpos[t])));
    Q4 q; q.w = L.quat[t][0]; q.x = L.quat[t][1]; q.y = L.quat[t][2]; q.z = L.quat[t][3];
    Q4 lq = qmul(hq, q);
    st3(ob + o_rot + 6 * b, qrot_ref(lq, v3(1.f, 0.f, 0.f)));
    st3(ob + o_rot + 6 * b + 3, qrot_ref(lq, v3(0.f, 0.f, 1.f)));
    if (K.cfg.self_obs_v == 2) {
      st3(ob + o_vel + 3 * b, qrot_ref(hq, ld3(&L.acc2[t][0])));
      st3(ob + o_vel + 3 * nb + 3 * b, qrot_ref(hq, ld3(&L.acc2[t][3])));
    } else {
      if (e.kind == CH_KIND_ROOTROT) st3(ob + o_vel + 3, qrot_ref(hq, ld3(L.qd[t])));
      else for (int k = 0; k < e.ndof; k++) ob[o_vel + e.dofadr + k] = L.qd[t][k];
    }
    if (e.kind == CH_KIND_ROOTROT) {
      int ot = K.self_obs_dim;
      if (K.cfg.task == SMPLSIM_TASK_SPEED) {
        V3 d = qrot_ref(hq, v3(1.f, 0.f, 0.f));
        ob[ot] = d.x; ob[ot + 1] = d.y; ob[ot + 2] = tk[0];
      } else if (K.cfg.task == SMPLSIM_TASK_REACH) st3(ob + ot, qrot_ref(hq, ld3(tk) - ld3(sc + CH_SC_ROOTPOS)));
      else if (K.cfg.task == SMPLSIM_TASK_GETUP) ob[ot] = tk[0];
    }
  }
}

template <int TMAX>
__device__ __noinline__ void ch_write_aux(const ChainConsts& K, const ChainCtx& x, ChainLoc<TMAX>& L, const float* sc, const SmplsimAux& aux, int env,
                                          const ChainFwd& fo) {
  if (!x.live) return;
  V3 root = ld3(sc + CH_SC_ROOTPOS);
  for (int t = 0; t < x.T; t++) {
    const ChainEntry& e = ch_entry(x, t);
    if (e.pb < 0) continue;
    if (aux.qacc) for (int k = 0; k < e.ndof; k++) aux.qacc[(size_t)env * K.nv + e.dofadr + k] = L.qacc[t][k];
    if (e.kind == CH_KIND_ROOTTRANS) continue;
    size_t bi = (size_t)env * K.nb + e.body;
    if (aux.xpos) st3(aux.xpos + bi * 3, ld3(L.xpos[t]) + root);
    if (aux.xquat) for (int j = 0; j < 4; j++) aux.xquat[bi * 4 + j] = L.quat[t][j];
    if (aux.body_linvel) st3(aux.body_linvel + bi * 3, ld3(&L.acc2[t][0]));
    if (aux.body_angvel) st3(aux.body_angvel + bi * 3, ld3(&L.acc2[t][3]));
    if (aux.ctrl && e.kind == CH_KIND_HINGE) for (int k = 0; k < e.ndof; k++) aux.ctrl[(size_t)env * K.nu + e.dofadr - 6 + k] = L.tau[t][k];
  }
  if (x.c == 0) {
    if (aux.contact_mask) aux.contact_mask[env] = fo.mask;
    if (aux.solver_iter) aux.solver_iter[env] = fo.iters;
  }
}

extern __shared__ float ch_smem[];

template <int TMAX>
__device__ __forceinline__ void ch_setup(const ChainEntry* gtab, const ChainConsts& K, ChainCtx& x, float*& sc, int& env, int n) {
  int ntab = K.T * CH_LPE * (int)(sizeof(ChainEntry) / 4);
  const int* src = (const int*)gtab;
  int* dst = (int*)ch_smem;
  for (int i = threadIdx.x; i < ntab; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  int e = lane >> 2;
  env = (blockIdx.x * wpb + wib) * CH_EPW + e;
  x.tab = (const ChainEntry*)ch_smem;
  x.c = lane & 3;
  x.T = K.T;
  x.live = env < n;
  float* envblk = ch_smem + ntab + (size_t)(wib * CH_EPW + e) * K.mb_stride;
  x.mb = envblk;
  sc = envblk + 19 * K.n_mbox + 27 * K.n_xedge;
}

template <int TMAX>
__global__ void __launch_bounds__(128) k_step2(const ChainEntry* __restrict__ gtab, ChainConsts K, ChainStepArgs a) {
  ChainCtx x; float* sc; int env;
  ch_setup<TMAX>(gtab, K, x, sc, env, a.n);
  int lane = threadIdx.x & 31;
  if (!__any_sync(CH_FULL, x.live)) return;
  ChainLoc<TMAX> L;
  for (int t = 0; t < TMAX; t++) L.lflag[t] = 0;
  const bool spd = (K.cfg.control_mode == SMPLSIM_CTRL_UHC_PD);
  size_t eo = x.live ? (size_t)env : 0;
  if (spd && K.cfg.spd_stale && a.mode == 0) {   // mj_data.qM / qfrc_bias as left by the last forward pass (quirk Q1)
    ch_load(K, x, L, a.st.qpos_fwd + eo * K.nq, a.st.qvel_fwd + eo * K.nv, nullptr, nullptr, 1, sc);
    __syncwarp();
    ch_fk(K, x, L, true);
    ch_spd_prepare(K, x, L);
  }
  ch_load(K, x, L, a.st.qpos + eo * K.nq, a.st.qvel + eo * K.nv, a.st.qacc_warm + eo * K.nv, a.action + eo * K.nu, 1 | 2 | (a.mode == 0 ? 4 : 8), sc);
  ch_task_io(x, sc, a.st, env, false);
  if (a.mode == 0 && x.live && x.c == 0) {
    int* ti = (int*)(sc + CH_SC_TSK);
    if (K.cfg.task != SMPLSIM_TASK_NONE && ti[5] >= ti[4]) ch_reset_task(K, sc, env);
  }
  __syncwarp();
  ChainFwd fo; fo.mask = 0ull; fo.iters = 0;
  V3 disp = ch_substeps(K, x, L, sc, lane, a.nsub, a.mode, &fo, a.st, env, true, false);
  ch_fk(K, x, L, false);   // mj_kinematics at the integrated state
  if (a.mode == 0) {
    int* ti = (int*)(sc + CH_SC_TSK);
    if (x.live && x.c == 0) ti[5] += 1;
    __syncwarp();
    ch_write_obs(K, x, L, sc, a.obs ? a.obs + eo * K.obs_dim : nullptr);
    if (x.live) {
      const SmplsimEnvCfg& c = K.cfg;
      const float* tk = sc + CH_SC_TSK;
      for (int t = 0; t < x.T; t++) {
        const ChainEntry& e = ch_entry(x, t);
        if (e.pb < 0) continue;
        if (e.kind == CH_KIND_ROOTTRANS) {
          float rew = 0.f;
          bool mine = true;
          if (c.task == SMPLSIM_TASK_SPEED) {
            float inv_dt = 1.0f / (K.h * (float)a.nsub), vx = disp.x * inv_dt, vy = disp.y * inv_dt, er = tk[0] - vx;
            rew = expf(-0.25f * (er * er + 0.1f * vy * vy));
          } else if (c.task == SMPLSIM_TASK_GETUP) { float er = tk[0] - sc[CH_SC_ROOTPOS + 2]; rew = expf(-4.0f * er * er); }
          else if (c.task == SMPLSIM_TASK_REACH) mine = false;
          if (mine && a.reward) a.reward[env] = rew;
        } else if (c.task == SMPLSIM_TASK_REACH && e.body == c.reach_body) {
          V3 dl = ld3(tk) - (ld3(L.xpos[t]) + ld3(sc + CH_SC_ROOTPOS));
          if (a.reward) a.reward[env] = expf(-4.0f * dot(dl, dl));
        }
      }
      if (x.c == 0) {
        int term = 0, trunc = 0, pass_time = ti[5] > c.episode_length;
        if (c.task == SMPLSIM_TASK_NONE) trunc = pass_time;
        else if (c.task == SMPLSIM_TASK_GETUP && ti[6] > 0) ti[6] -= 1;
        else { trunc = pass_time; term = (fo.mask & ~K.legal_mask) != 0ull; }
        if (a.terminated) a.terminated[env] = (uint8_t)term;
        if (a.truncated) a.truncated[env] = (uint8_t)trunc;
      }
    }
  }
  ch_write_aux(K, x, L, sc, a.aux, env, fo);
  ch_store(K, x, L, a.st.qpos + eo * K.nq, a.st.qvel + eo * K.nv, a.st.qacc_warm + eo * K.nv);
  __syncwarp();
  if (a.mode == 0) ch_task_io(x, sc, a.st, env, true);
}

template <int TMAX>
__global__ void __launch_bounds__(128) k_reset2(const ChainEntry* __restrict__ gtab, ChainConsts K, ChainResetArgs a) {
  ChainCtx x; float* sc; int env;
  ch_setup<TMAX>(gtab, K, x, sc, env, a.n);
  int lane = threadIdx.x & 31;
  if (x.live && a.mask && !a.mask[env]) x.live = false;
  if (!__any_sync(CH_FULL, x.live)) return;
  ChainLoc<TMAX> L;
  const SmplsimEnvCfg& c = K.cfg;
  size_t eo = x.live ? (size_t)env : 0;
  int init = a.init_mode < 0 ? c.state_init : a.init_mode;
  ch_task_io(x, sc, a.st, env, false);
  if (x.live && x.c == 0) {
    int* ti = (int*)(sc + CH_SC_TSK);
    if (c.task == SMPLSIM_TASK_GETUP) ti[6] = c.recovery_steps;
    if (!c.legacy_change_step) ti[5] = 0;
    ch_reset_task(K, sc, env);   // sees the old cur_t when legacy_change_step (quirk Q4)
  }
  for (int t = 0; t < TMAX; t++) {
    L.lflag[t] = 0;
    for (int k = 0; k < 3; k++) { L.q[t][k] = 0.f; L.qd[t][k] = 0.f; L.qacc[t][k] = 0.f; L.tau[t][k] = 0.f; L.act[t][k] = 0.f; }
  }
  L.rootq[0] = 1.f; L.rootq[1] = L.rootq[2] = L.rootq[3] = 0.f;
  ChainFwd fo; fo.mask = 0ull; fo.iters = 0;
  if (init == SMPLSIM_INIT_MOCAP) {
    ch_load(K, x, L, a.qpos0 + eo * K.nq, a.qvel0 + eo * K.nv, nullptr, nullptr, 1, sc);
  } else if (x.live) {
    float z = (init == SMPLSIM_INIT_DEFAULT) ? 0.94f : 0.3f;
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb < 0) continue;
      if (e.kind == CH_KIND_ROOTTRANS) { L.q[t][2] = z; st3(sc + CH_SC_ROOTPOS, v3(0.f, 0.f, z)); }
      if (e.kind == CH_KIND_ROOTROT && init == SMPLSIM_INIT_DEFAULT) { L.rootq[0] = L.rootq[1] = L.rootq[2] = L.rootq[3] = 0.5f; }
    }
  }
  __syncwarp();
  if (init == SMPLSIM_INIT_FALL) {
    if (c.control_mode == SMPLSIM_CTRL_UHC_PD && c.spd_stale) { ch_fk(K, x, L, true); ch_spd_prepare(K, x, L); }   // mj_forward
    int ngrp = (K.nu + 3) / 4;
    for (int k3 = 0; k3 < 3; k3++) {
      int* ti = (int*)(sc + CH_SC_TSK);
      uint32_t base = x.live ? (uint32_t)ti[7] : 0u;
      if (x.live) {
        for (int t = 0; t < x.T; t++) {
          const ChainEntry& e = ch_entry(x, t);
          if (e.pb < 0 || e.kind != CH_KIND_HINGE) continue;
          for (int k = 0; k < e.ndof; k++) {
            int i = e.dofadr - 6 + k;
            uint32_t r[4];
            philox4x32(base + (uint32_t)(i >> 2), (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
            L.act[t][k] = u01(r[i & 3]) - 0.5f;
          }
        }
      }
      __syncwarp();
      if (x.live && x.c == 0) ti[7] = (int)(base + (uint32_t)ngrp);
      __syncwarp();
      ch_substeps(K, x, L, sc, lane, c.nsubsteps, 0, &fo, a.st, env, false, true);
    }
  }
  // reset_sim(): mj_forward at the reset state
  ch_fk(K, x, L, true);
  {
    float root_h = dot(ld3(K.plane_n), ld3(sc + CH_SC_ROOTPOS) - ld3(K.plane_pos));
    int nrows = 0;
    unsigned long long m = ch_collide(K, x, L, root_h, &nrows);
    unsigned lo = (unsigned)(m & 0xffffffffull), hi = (unsigned)(m >> 32);
    lo |= __shfl_xor_sync(CH_FULL, lo, 1); lo |= __shfl_xor_sync(CH_FULL, lo, 2);
    hi |= __shfl_xor_sync(CH_FULL, hi, 1); hi |= __shfl_xor_sync(CH_FULL, hi, 2);
    fo.mask = ((unsigned long long)hi << 32) | lo;
  }
  if (x.live) {
    for (int t = 0; t < x.T; t++) {
      const ChainEntry& e = ch_entry(x, t);
      if (e.pb < 0) continue;
      S6 v = ld6(L.vel[t]);
      st3(&L.acc2[t][0], v.l + cross(v.a, ld3(L.xpos[t])));
      st3(&L.acc2[t][3], v.a);
    }
    if (x.c == 0) ((int*)(sc + CH_SC_TSK))[5] = 0;
  }
  __syncwarp();
  ch_write_obs(K, x, L, sc, a.obs ? a.obs + eo * K.obs_dim : nullptr);
  ch_write_aux(K, x, L, sc, a.aux, env, fo);
  ch_store(K, x, L, a.st.qpos + eo * K.nq, a.st.qvel + eo * K.nv, a.st.qacc_warm + eo * K.nv);
  ch_store(K, x, L, a.st.qpos_fwd + eo * K.nq, a.st.qvel_fwd + eo * K.nv, nullptr);
  __syncwarp();
  ch_task_io(x, sc, a.st, env, true);
}
