// physics.cuh -- one physics substep for one env per warp (sm_100a).
//
// Replaces, per env and per substep, ctrler.control(...) + mujoco.mj_step(...)
// (smpl_sim/envs/humanoid_env.py:444-450).  NOT MuJoCo's algorithm: forward dynamics and the
// soft-constraint contact solve are done with the Articulated-Body Algorithm in world-aligned
// coordinates about the root origin, so no mass matrix is ever formed or factorised:
//
//   * (M + h Kd)^-1 of the stable-PD controller  = ABA with armature + h*kd      (controllers.py:165-190)
//   * qacc = argmin 1/2|a-a0|_M^2 + sum 1/2 D_i min(0, J_i a - aref_i)^2          (MuJoCo's convex problem)
//       for floor contacts every row touches ONE body, so for a fixed active set the minimiser is
//       an ABA whose articulated inertia of that body gains sum D_i x_i x_i^T and whose bias force
//       gains -sum D_i aref_i x_i  (x_i = unit contact wrench).  Active-set Newton with an exact
//       line search carried out entirely in constraint-row space (no M products needed).
//
// Mapping: one env per warp; tree sweeps are level-synchronous (bodies of one depth in parallel
// lanes, __syncwarp between depths); per-body spatial inertias, joint transforms and the ABA
// factors U, 1/D live in this warp's shared-memory scratch (EnvLayout).
#pragma once
#include "dev_model.cuh"

#define FULLMASK 0xffffffffu
#define SOLVER_MAXITER 12
#define LS_MAXITER 24

// ------------------------------------------------------------------ motion subspace of dof d (about the root origin)
__device__ __forceinline__ S6 dof_S(const DevModel& M, int b, int k, const float* ax, const float* xpos) {
  int d = M.dofadr[b] + k;
  V3 a = ld3(ax + 3 * d);
  if (b == 0 && k < 3) return s6(v3(0.f, 0.f, 0.f), a);
  return s6(a, cross(ld3(xpos + 3 * b), a));
}

// ------------------------------------------------------------------ kinematics (+ velocities, bias forces)   mj_kinematics / comVel / rne
template <bool VEL>
__device__ void fk_pass(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  float *qpos = sm + L.qpos, *qvel = sm + L.qvel, *xpos = sm + L.xpos, *xquat = sm + L.xquat, *xmat = sm + L.xmat;
  float *ax = sm + L.ax, *vel = sm + L.vel, *abias = sm + L.abias, *pb = sm + L.pb, *irb = sm + L.irb;
  for (int lev = 0; lev < M.nlevel; lev++) {
    for (int i = M.level_adr[lev] + lane; i < M.level_adr[lev + 1]; i += 32) {
      int b = M.level_list[i];
      Q4 qc;
      V3 x;
      S6 v, ab;
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
        if (VEL) {
          V3 vl = ld3(qvel), w = qvel[3] * c0 + qvel[4] * c1 + qvel[5] * c2;
          v = s6(w, vl);
          ab = s6(v3(0.f, 0.f, 0.f), v3(-M.grav[0], -M.grav[1], -M.grav[2]) + cross(vl, w));
        }
      } else {
        int p = M.parent[b];
        x = ld3(xpos + 3 * p) + mrot(xmat + 9 * p, ld3(M.bpos[b]));
        Q4 qp, qb;
        qp.w = xquat[4 * p]; qp.x = xquat[4 * p + 1]; qp.y = xquat[4 * p + 2]; qp.z = xquat[4 * p + 3];
        qb.w = M.bquat[b][0]; qb.x = M.bquat[b][1]; qb.y = M.bquat[b][2]; qb.z = M.bquat[b][3];
        qc = qmul(qp, qb);
        if (VEL) { v = ld6(vel + 6 * p); ab = ld6(abias + 6 * p); }
        int d0 = M.dofadr[b], nd = M.dofnum[b];
        for (int k = 0; k < nd; k++) {
          int d = d0 + k;
          V3 al = ld3(M.axis[d]);
          V3 a = qrot(qc, al);
          st3(ax + 3 * d, a);
          if (VEL) {
            S6 S = s6(a, cross(x, a));
            float qd = qvel[d];
            S6 sd = cross_motion(v, S);
            ab = ab + qd * sd;
            v = v + qd * S;
          }
          float sn, cs;
          sincosf(0.5f * qpos[d + 1], &sn, &cs);
          Q4 qj;
          qj.w = cs; qj.x = al.x * sn; qj.y = al.y * sn; qj.z = al.z * sn;
          qc = qmul(qc, qj);
        }
        qc = qnormalize(qc);
      }
      float R[9];
      q2mat(qc, R);
      xquat[4 * b] = qc.w; xquat[4 * b + 1] = qc.x; xquat[4 * b + 2] = qc.y; xquat[4 * b + 3] = qc.z;
#pragma unroll
      for (int j = 0; j < 9; j++) xmat[9 * b + j] = R[j];
      st3(xpos + 3 * b, x);
      if (VEL) {
        st6(vel + 6 * b, v);
        st6(abias + 6 * b, ab);
        // rigid inertia about the root origin, world axes
        const float* in = M.inertia[b];
        float m = M.mass[b];
        V3 r = x + mrot(R, ld3(M.ipos[b]));
        float Il[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, T[9];
#pragma unroll
        for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
          for (int j = 0; j < 3; j++) T[3 * i2 + j] = R[3 * i2] * Il[j] + R[3 * i2 + 1] * Il[3 + j] + R[3 * i2 + 2] * Il[6 + j];
        float rr = dot(r, r);
        float r10[10];
        r10[0] = m; r10[1] = m * r.x; r10[2] = m * r.y; r10[3] = m * r.z;
        r10[4] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + m * (rr - r.x * r.x);
        r10[5] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + m * (rr - r.y * r.y);
        r10[6] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + m * (rr - r.z * r.z);
        r10[7] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - m * r.x * r.y;
        r10[8] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - m * r.x * r.z;
        r10[9] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - m * r.y * r.z;
#pragma unroll
        for (int j = 0; j < 10; j++) irb[10 * b + j] = r10[j];
        S6 f = rb_mul(r10, ab) + cross_force(v, rb_mul(r10, v));
        st6(pb + 6 * b, f);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ contact row helpers
__device__ __forceinline__ S6 contact_wrench(const DevModel& M, V3 cp, V3 t1, int k) {
  V3 n = ld3(M.plane_n);
  V3 t = (k < 2) ? t1 : cross(n, t1);
  float sg = (k & 1) ? -M.mu : M.mu;
  V3 dir = n + sg * t;
  return s6(cross(cp, dir), dir);
}

// ------------------------------------------------------------------ ABA inward sweep (inertia and/or force)
// INERTIA: builds U, 1/D per dof from rigid inertias (+ active contact rows, + dadd on the joint diagonal).
// FORCE:   u = tin - S.pA, pA += U u / D, with pA initialised from the bias force pb (PB) and contact terms.
template <bool INERTIA, bool FORCE, bool PB, bool CONTACTS>
__device__ void aba_inward(const DevModel& M, const EnvLayout& L, float* sm, int lane, const float* ax, const float* xpos,
                           float* Uarr, float* Dinv, const float* dadd, const float* tin) {
  float *IA = sm + L.IA, *pA = sm + L.pA, *u = sm + L.u;
  const float *irb = sm + L.irb, *pb = sm + L.pb;
  for (int lev = M.nlevel - 1; lev >= 0; lev--) {
    for (int i = M.level_adr[lev] + lane; i < M.level_adr[lev + 1]; i += 32) {
      int b = M.level_list[i];
      float A[21];
      S6 p = s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f));
      if (INERTIA) rb_expand(irb + 10 * b, A);
      if (FORCE && PB) p = ld6(pb + 6 * b);
      if (CONTACTS) {
        const float *cpos = sm + L.cpos, *ct1 = sm + L.ct1, *cD = sm + L.cD, *caref = sm + L.caref;
        const int* cflag = (const int*)(sm + L.cflag);
        for (int gi = M.bgeom_adr[b]; gi < M.bgeom_adr[b + 1]; gi++) {
          int g = M.bgeom_list[gi];
          for (int c = M.slot_adr[g]; c < M.slot_adr[g + 1]; c++) {
            int fl = cflag[c];
            if (!(fl & 1)) continue;
            V3 cp = ld3(cpos + 3 * c), t1 = ld3(ct1 + 3 * c);
            float D = cD[c];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (!(fl & (2 << k))) continue;
              S6 xw = contact_wrench(M, cp, t1, k);
              float xv[6] = {xw.a.x, xw.a.y, xw.a.z, xw.l.x, xw.l.y, xw.l.z};
              if (INERTIA) sym_rank1(A, xv, -D);
              if (FORCE) p = p - (D * caref[4 * c + k]) * xw;
            }
          }
        }
      }
      for (int ci = M.child_adr[b]; ci < M.child_adr[b + 1]; ci++) {
        int c = M.child_list[ci];
        if (INERTIA) {
#pragma unroll
          for (int j = 0; j < 21; j++) A[j] += IA[21 * c + j];
        }
        if (FORCE) p = p + ld6(pA + 6 * c);
      }
      int d0 = M.dofadr[b];
      for (int k = M.dofnum[b] - 1; k >= 0; k--) {
        int d = d0 + k;
        S6 S = dof_S(M, b, k, ax, xpos);
        float s[6] = {S.a.x, S.a.y, S.a.z, S.l.x, S.l.y, S.l.z}, Uv[6], di;
        if (INERTIA) {
          sym_mul(A, s, Uv);
          float D = M.arm[d] + (dadd ? dadd[d] : 0.f);
#pragma unroll
          for (int j = 0; j < 6; j++) D = fmaf(s[j], Uv[j], D);
          di = 1.0f / D;
#pragma unroll
          for (int j = 0; j < 6; j++) Uarr[6 * d + j] = Uv[j];
          Dinv[d] = di;
          sym_rank1(A, Uv, di);
        } else {
#pragma unroll
          for (int j = 0; j < 6; j++) Uv[j] = Uarr[6 * d + j];
          di = Dinv[d];
        }
        if (FORCE) {
          float uu = tin[d] - dot6(S, p);
          u[d] = uu;
          float c = uu * di;
          p = p + c * s6(v3(Uv[0], Uv[1], Uv[2]), v3(Uv[3], Uv[4], Uv[5]));
        }
      }
      if (INERTIA) {
#pragma unroll
        for (int j = 0; j < 21; j++) IA[21 * b + j] = A[j];
      }
      if (FORCE) st6(pA + 6 * b, p);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ABA outward sweep: qdd = (u - U.a)/D, a += S qdd
__device__ void aba_outward(const DevModel& M, const EnvLayout& L, float* sm, int lane, const float* ax, const float* xpos,
                            const float* Uarr, const float* Dinv, float* qout) {
  float *acc = sm + L.acc, *u = sm + L.u;
  for (int lev = 0; lev < M.nlevel; lev++) {
    for (int i = M.level_adr[lev] + lane; i < M.level_adr[lev + 1]; i += 32) {
      int b = M.level_list[i];
      S6 a = (b == 0) ? s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)) : ld6(acc + 6 * M.parent[b]);
      int d0 = M.dofadr[b], nd = M.dofnum[b];
      for (int k = 0; k < nd; k++) {
        int d = d0 + k;
        S6 S = dof_S(M, b, k, ax, xpos);
        S6 Uv = ld6(Uarr + 6 * d);
        float qdd = Dinv[d] * (u[d] - dot6(Uv, a));
        a = a + qdd * S;
        qout[d] = qdd;
      }
      st6(acc + 6 * b, a);
    }
    __syncwarp();
  }
}

// spatial "J qacc" acceleration of every body for a given qacc (warm start)
__device__ void acc_from_qacc(const DevModel& M, const EnvLayout& L, float* sm, int lane, const float* q) {
  float *acc = sm + L.acc;
  const float *ax = sm + L.ax, *xpos = sm + L.xpos;
  for (int lev = 0; lev < M.nlevel; lev++) {
    for (int i = M.level_adr[lev] + lane; i < M.level_adr[lev + 1]; i += 32) {
      int b = M.level_list[i];
      S6 a = (b == 0) ? s6(v3(0.f, 0.f, 0.f), v3(0.f, 0.f, 0.f)) : ld6(acc + 6 * M.parent[b]);
      int d0 = M.dofadr[b], nd = M.dofnum[b];
      for (int k = 0; k < nd; k++) a = a + q[d0 + k] * dof_S(M, b, k, ax, xpos);
      st6(acc + 6 * b, a);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ collision: floor plane vs capsule / sphere / box   (mj_collision)
__device__ __forceinline__ float impedance(const DevModel& M, float pm) {
  float x = fabsf(pm) / fmaxf(M.solimp[2], 1e-15f);
  if (x >= 1.f) return M.solimp[1];
  if (x <= 0.f) return M.solimp[0];
  float y, pw = M.solimp[4];
  if (pw < 1.0000001f && pw > 0.9999999f) y = x;
  else if (x <= M.solimp[3]) y = M.imp_a * powf(x, pw);
  else y = 1.f - M.imp_b * powf(1.f - x, pw);
  return M.solimp[0] + y * (M.solimp[1] - M.solimp[0]);
}

__device__ __forceinline__ void emit_contact(const DevModel& M, const EnvLayout& L, float* sm, int c, int b, V3 cp, float dist, V3 hint, bool has_hint) {
  V3 n = ld3(M.plane_n), t1;
  if (has_hint) {
    t1 = hint - dot(n, hint) * n;
    float nn = sqrtf(dot(t1, t1));
    t1 = (nn < 1e-15f) ? v3(1.f, 0.f, 0.f) : (1.0f / nn) * t1;
  } else t1 = ld3(M.t1_default);
  st3(sm + L.cpos + 3 * c, cp);
  st3(sm + L.ct1 + 3 * c, t1);
  float pm = dist - M.margin, imp = impedance(M, pm);
  float tran = M.tran_iw0[b], mu0 = M.mu;
  float R0 = fmaxf((1.f - imp) / imp * (tran + mu0 * mu0 * tran), 1e-15f);
  float R1 = R0 / fmaxf(M.impratio, 1e-15f);
  float mu = mu0 * sqrtf(R1 / R0);
  sm[L.cD + c] = 1.0f / (2.f * mu * mu * R0);
  S6 v = ld6(sm + L.vel + 6 * b);
  float kterm = M.K * imp * pm;
#pragma unroll
  for (int k = 0; k < 4; k++) sm[L.caref + 4 * c + k] = -M.B * dot6(contact_wrench(M, cp, t1, k), v) - kterm;
  ((int*)(sm + L.cflag))[c] = 1;
}

__device__ unsigned long long collide(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  const float *xpos = sm + L.xpos, *xmat = sm + L.xmat, *qpos = sm + L.qpos;
  int* cflag = (int*)(sm + L.cflag);
  V3 n = ld3(M.plane_n);
  float h0 = dot(n, ld3(qpos) - ld3(M.plane_pos));
  unsigned long long mask = 0ull;
  for (int g = lane; g < M.ng; g += 32) {
    int b = M.gbody[g];
    const float* R = xmat + 9 * b;
    V3 c = ld3(xpos + 3 * b) + mrot(R, ld3(M.gpos[g]));
    float d0 = h0 + dot(n, c);
    int s0 = M.slot_adr[g], s1 = M.slot_adr[g + 1], cnt = 0;
    for (int s = s0; s < s1; s++) cflag[s] = 0;
    const float* gm = M.gmat[g];
    int ty = M.gtype[g];
    if (ty == SMPLSIM_GEOM_CAPSULE || ty == SMPLSIM_GEOM_SPHERE) {
      V3 axl = v3(gm[2], gm[5], gm[8]);
      V3 axw = mrot(R, axl);
      float rad = M.gsize[g][0], hl = (ty == SMPLSIM_GEOM_CAPSULE) ? M.gsize[g][1] : 0.f;
      int nend = (ty == SMPLSIM_GEOM_CAPSULE) ? 2 : 1;
      float na = dot(n, axw);
      for (int e = 0; e < nend; e++) {
        float sg = e ? -hl : hl;
        float dist = d0 + sg * na - rad;
        if (dist > M.margin) continue;
        V3 p = c + sg * axw;
        emit_contact(M, L, sm, s0 + cnt, b, p - (rad + 0.5f * dist) * n, dist, axw, nend == 2);
        cnt++;
      }
    } else if (ty == SMPLSIM_GEOM_BOX) {
      for (int i = 0; i < 8 && cnt < 4; i++) {
        V3 vl = v3((i & 1) ? M.gsize[g][0] : -M.gsize[g][0], (i & 2) ? M.gsize[g][1] : -M.gsize[g][1], (i & 4) ? M.gsize[g][2] : -M.gsize[g][2]);
        V3 w = mrot(R, mrot(gm, vl));
        float l = dot(n, w);
        if (d0 + l > M.margin || l > 0.f) continue;
        float dist = d0 + l;
        emit_contact(M, L, sm, s0 + cnt, b, c + w - (0.5f * dist) * n, dist, n, false);
        cnt++;
      }
    }
    if (cnt) mask |= 1ull << (g + 1);
  }
  unsigned lo = __reduce_or_sync(FULLMASK, (unsigned)(mask & 0xffffffffull));
  unsigned hi = __reduce_or_sync(FULLMASK, (unsigned)(mask >> 32));
  __syncwarp();
  return ((unsigned long long)hi << 32) | lo;
}

// joint-limit rows (mj_instantiateLimit, margin 0): returns number of active limit rows in the warp
__device__ int make_limits(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  const float *qpos = sm + L.qpos, *qvel = sm + L.qvel;
  int* lflag = (int*)(sm + L.lflag);
  int n = 0;
  for (int d = 6 + lane; d < M.nv; d += 32) {
    int fl = 0;
    if (M.limited[d]) {
      float q = qpos[d + 1];
      float dlo = q - M.range[d][0], dhi = M.range[d][1] - q;
      float dist = 0.f, sg = 0.f;
      if (dlo < 0.f) { dist = dlo; sg = 1.f; fl = 1; }
      else if (dhi < 0.f) { dist = dhi; sg = -1.f; fl = 2; }
      if (fl) {
        float imp = impedance(M, dist);
        float R = fmaxf((1.f - imp) / imp * M.diw0[d], 1e-15f);
        sm[L.lD + d] = 1.0f / R;
        sm[L.laref + d] = -M.B * sg * qvel[d] - M.K * imp * dist;
        n++;
      }
    }
    lflag[d] = fl;  // bit0|bit1: which side; bit 2 set later when the row is active in the working set
  }
  n = __reduce_add_sync(FULLMASK, n);
  __syncwarp();
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}

// ------------------------------------------------------------------ constraint rows: evaluation at a body-acceleration field
// mode 0: set the working set from the row values (r<0) ; mode 1: store rs, report whether (rs<0) equals the working set
template <int MODE>
__device__ bool eval_rows(const DevModel& M, const EnvLayout& L, float* sm, int lane, const float* qd) {
  const float *acc = sm + L.acc, *cpos = sm + L.cpos, *ct1 = sm + L.ct1, *caref = sm + L.caref;
  int *cflag = (int*)(sm + L.cflag), *lflag = (int*)(sm + L.lflag);
  bool same = true;
  for (int c = lane; c < M.nslot; c += 32) {
    int fl = cflag[c];
    if (!(fl & 1)) continue;
    int b = M.gbody[M.slot_geom[c]];
    S6 a = ld6(acc + 6 * b);
    V3 cp = ld3(cpos + 3 * c), t1 = ld3(ct1 + 3 * c);
    int nf = 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float r = dot6(contact_wrench(M, cp, t1, k), a) - caref[4 * c + k];
      if (MODE == 1) sm[L.crs + 4 * c + k] = r;
      if (r < 0.f) nf |= 2 << k;
    }
    if (MODE == 0) cflag[c] = nf;
    else if (nf != fl) same = false;
  }
  for (int d = 6 + lane; d < M.nv; d += 32) {
    int fl = lflag[d];
    if (!(fl & 3)) continue;
    float sg = (fl & 1) ? 1.f : -1.f;
    float r = sg * qd[d] - sm[L.laref + d];
    if (MODE == 1) sm[L.lrs + d] = r;
    int nf = (fl & 3) | ((r < 0.f) ? 4 : 0);
    if (MODE == 0) lflag[d] = nf;
    else if (nf != fl) same = false;
  }
  same = __all_sync(FULLMASK, same);
  __syncwarp();
  return same;
}

// line-search sums over this lane's rows: s1 = sum_{r+al*d<0} D (r+al*d) d,  s2 = sum_{...} D d^2
__device__ __forceinline__ void ls_sums(const DevModel& M, const EnvLayout& L, const float* sm, int lane, float al, float& s1, float& s2) {
  const int *cflag = (const int*)(sm + L.cflag), *lflag = (const int*)(sm + L.lflag);
  s1 = 0.f; s2 = 0.f;
  for (int c = lane; c < M.nslot; c += 32) {
    if (!(cflag[c] & 1)) continue;
    float D = sm[L.cD + c];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float r = sm[L.cr + 4 * c + k], d = sm[L.crs + 4 * c + k] - r, v = fmaf(al, d, r);
      if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
    }
  }
  for (int dd = 6 + lane; dd < M.nv; dd += 32) {
    if (!(lflag[dd] & 3)) continue;
    float D = sm[L.lD + dd], r = sm[L.lr + dd], d = sm[L.lrs + dd] - r, v = fmaf(al, d, r);
    if (v < 0.f) { s1 = fmaf(D * v, d, s1); s2 = fmaf(D * d, d, s2); }
  }
}

// ------------------------------------------------------------------ constraint solve (mj_fwdConstraint), ABA formulation
// in: tau[nu], contact slots / limit rows prepared at the current state, qwarm.  out: qacc.  returns #ABA solves - 1.
__device__ int solve_constrained(const DevModel& M, const EnvLayout& L, float* sm, int lane, bool any_rows) {
  float *tin = sm + L.tin, *tau = sm + L.tau, *qacc = sm + L.qacc, *qstar = sm + L.qstar, *dadd = sm + L.dadd;
  const float *ax = sm + L.ax, *xpos = sm + L.xpos;
  float *U = sm + L.U, *Dinv = sm + L.Dinv;
  int *cflag = (int*)(sm + L.cflag), *lflag = (int*)(sm + L.lflag);
  for (int d = lane; d < M.nv; d += 32) { tin[d] = d < 6 ? 0.f : tau[d - 6]; dadd[d] = 0.f; }
  __syncwarp();
  if (!any_rows) {
    aba_inward<true, true, true, false>(M, L, sm, lane, ax, xpos, U, Dinv, nullptr, tin);
    aba_outward(M, L, sm, lane, ax, xpos, U, Dinv, qacc);
    return 0;
  }
  acc_from_qacc(M, L, sm, lane, sm + L.qwarm);
  eval_rows<0>(M, L, sm, lane, sm + L.qwarm);
  bool have_point = false;
  int it = 0;
  for (; it < SOLVER_MAXITER; it++) {
    for (int d = 6 + lane; d < M.nv; d += 32) {
      int fl = lflag[d];
      float t = tau[d - 6], da = 0.f;
      if (fl & 4) { da = sm[L.lD + d]; t += ((fl & 1) ? 1.f : -1.f) * da * sm[L.laref + d]; }
      tin[d] = t; dadd[d] = da;
    }
    __syncwarp();
    aba_inward<true, true, true, true>(M, L, sm, lane, ax, xpos, U, Dinv, dadd, tin);
    aba_outward(M, L, sm, lane, ax, xpos, U, Dinv, qstar);
    bool same = eval_rows<1>(M, L, sm, lane, qstar);
    if (same) {
      for (int d = lane; d < M.nv; d += 32) qacc[d] = qstar[d];
      __syncwarp();
      return it;
    }
    if (!have_point) {
      // adopt a* as the current iterate: M a* - qfrc_smooth = J_A^T (-D rs)_A  with A the working set just used
      for (int d = lane; d < M.nv; d += 32) qacc[d] = qstar[d];
      for (int c = lane; c < M.nslot; c += 32) {
        int fl = cflag[c];
        if (!(fl & 1)) continue;
        float D = sm[L.cD + c];
        int nf = 1;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          float rs = sm[L.crs + 4 * c + k];
          sm[L.cr + 4 * c + k] = rs;
          sm[L.cphi + 4 * c + k] = (fl & (2 << k)) ? -D * rs : 0.f;
          if (rs < 0.f) nf |= 2 << k;
        }
        cflag[c] = nf;
      }
      for (int d = 6 + lane; d < M.nv; d += 32) {
        int fl = lflag[d];
        if (!(fl & 3)) continue;
        float rs = sm[L.lrs + d];
        sm[L.lr + d] = rs;
        sm[L.lphi + d] = (fl & 4) ? -sm[L.lD + d] * rs : 0.f;
        lflag[d] = (fl & 3) | ((rs < 0.f) ? 4 : 0);
      }
      have_point = true;
      __syncwarp();
      continue;
    }
    // exact line search on psi(al) = al g1 + al^2/2 g2 + sum 1/2 D min(0, r + al d)^2   (row space only)
    float g1 = 0.f, g2 = 0.f;
    for (int c = lane; c < M.nslot; c += 32) {
      int fl = cflag[c];
      if (!(fl & 1)) continue;
      float D = sm[L.cD + c];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float r = sm[L.cr + 4 * c + k], rs = sm[L.crs + 4 * c + k], ph = sm[L.cphi + 4 * c + k];
        float phs = (fl & (2 << k)) ? -D * rs : 0.f, d = rs - r;
        g1 = fmaf(d, ph, g1); g2 = fmaf(d, phs - ph, g2);
      }
    }
    for (int d = 6 + lane; d < M.nv; d += 32) {
      int fl = lflag[d];
      if (!(fl & 3)) continue;
      float r = sm[L.lr + d], rs = sm[L.lrs + d], ph = sm[L.lphi + d];
      float phs = (fl & 4) ? -sm[L.lD + d] * rs : 0.f, dl = rs - r;
      g1 = fmaf(dl, ph, g1); g2 = fmaf(dl, phs - ph, g2);
    }
    g1 = warp_sum(g1); g2 = warp_sum(g2);
    float s1, s2;
    ls_sums(M, L, sm, lane, 0.f, s1, s2);
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    float f0 = g1 + s1, al = 0.f;
    if (f0 < -1e-4f * (fabsf(g1) + fabsf(s1))) {   // below the fp32 cancellation floor: converged
      float lo = 0.f, hi = -1.f, tol = 1e-6f * fabsf(f0);
      al = 1.f;
      for (int ls = 0; ls < LS_MAXITER; ls++) {
        ls_sums(M, L, sm, lane, al, s1, s2);
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        float f = g1 + al * g2 + s1, fp = g2 + s2;
        if (fabsf(f) <= tol) break;
        if (f < 0.f) lo = al; else hi = al;
        float an = (fp > 0.f) ? al - f / fp : -1.f;
        if (!(an > lo) || (hi > 0.f && !(an < hi))) an = (hi > 0.f) ? 0.5f * (lo + hi) : 2.f * al;
        an = fminf(an, 16.f);
        if (an == al) break;
        al = an;
      }
    }
    if (!(al > 0.f)) break;
    for (int d = lane; d < M.nv; d += 32) qacc[d] = fmaf(al, qstar[d] - qacc[d], qacc[d]);
    for (int c = lane; c < M.nslot; c += 32) {
      int fl = cflag[c];
      if (!(fl & 1)) continue;
      float D = sm[L.cD + c];
      int nf = 1;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float r = sm[L.cr + 4 * c + k], rs = sm[L.crs + 4 * c + k], ph = sm[L.cphi + 4 * c + k];
        float phs = (fl & (2 << k)) ? -D * rs : 0.f;
        r = fmaf(al, rs - r, r);
        sm[L.cr + 4 * c + k] = r;
        sm[L.cphi + 4 * c + k] = fmaf(al, phs - ph, ph);
        if (r < 0.f) nf |= 2 << k;
      }
      cflag[c] = nf;
    }
    for (int d = 6 + lane; d < M.nv; d += 32) {
      int fl = lflag[d];
      if (!(fl & 3)) continue;
      float r = sm[L.lr + d], rs = sm[L.lrs + d], ph = sm[L.lphi + d];
      float phs = (fl & 4) ? -sm[L.lD + d] * rs : 0.f;
      r = fmaf(al, rs - r, r);
      sm[L.lr + d] = r;
      sm[L.lphi + d] = fmaf(al, phs - ph, ph);
      lflag[d] = (fl & 3) | ((r < 0.f) ? 4 : 0);
    }
    __syncwarp();
  }
  __syncwarp();
  return it;
}

// ------------------------------------------------------------------ stable PD: factors of (M + h Kd) at the current state   (controllers.py:165-190)
__device__ void spd_prepare(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  float *dadd = sm + L.dadd, *tin = sm + L.tin;
  for (int d = lane; d < M.nv; d += 32) { dadd[d] = d < 6 ? 0.f : M.h * M.kd[d - 6]; tin[d] = 0.f; }
  for (int i = lane; i < 3 * M.nv; i += 32) sm[L.spd_ax + i] = sm[L.ax + i];
  for (int i = lane; i < 3 * M.nb; i += 32) sm[L.spd_xpos + i] = sm[L.xpos + i];
  __syncwarp();
  // a_bias = (M + h Kd)^-1 (-C): bias force in, zero joint force
  aba_inward<true, true, true, false>(M, L, sm, lane, sm + L.spd_ax, sm + L.spd_xpos, sm + L.spd_U, sm + L.spd_Dinv, dadd, tin);
  aba_outward(M, L, sm, lane, sm + L.spd_ax, sm + L.spd_xpos, sm + L.spd_U, sm + L.spd_Dinv, sm + L.spd_ab);
}

// torque for this substep from action `act` (controllers.py:116-163 | 316-346 | 26-47) -> tau
__device__ void compute_torque(const DevModel& M, const EnvLayout& L, float* sm, int lane, const SmplsimState& st, int env) {
  const float *qpos = sm + L.qpos, *qvel = sm + L.qvel, *act = sm + L.act;
  float *tau = sm + L.tau, *tin = sm + L.tin;
  int mode = M.cfg.control_mode;
  if (mode == SMPLSIM_CTRL_TORQUE) {
    for (int i = lane; i < M.nu; i += 32) tau[i] = fminf(fmaxf(act[i] * M.ascale[i], -M.tlim[i]), M.tlim[i]);
    __syncwarp();
    return;
  }
  if (mode == SMPLSIM_CTRL_SIMPLE_PID) {
    float dt = M.h * (float)M.cfg.nsubsteps;
    for (int i = lane; i < M.nu; i += 32) {
      size_t o = (size_t)env * M.nu + i;
      float lim = M.tlim[i], err = fmaf(act[i], M.ascale[i], M.aoffset[i]) - qpos[7 + i], le = st.pid_last_error[o];
      float derr = (le != le) ? 0.f : err - le;
      float in = fminf(fmaxf(fmaf(err, dt, st.pid_integral[o]), -lim), lim);
      st.pid_integral[o] = in; st.pid_last_error[o] = err;
      tau[i] = fminf(fmaxf(M.kp[i] * err + in + M.kd[i] * derr / dt, -lim), lim);
    }
    __syncwarp();
    return;
  }
  if (mode == SMPLSIM_CTRL_PD) {
    for (int i = lane; i < M.nu; i += 32) {
      float tgt = fmaf(act[i], M.ascale[i], M.aoffset[i]);
      float t = -M.kp[i] * (qpos[7 + i] - tgt) - M.kd[i] * qvel[6 + i];
      tau[i] = fminf(fmaxf(t, -M.tlim[i]), M.tlim[i]);
    }
    __syncwarp();
    return;
  }
  for (int d = lane; d < M.nv; d += 32) {
    float t = 0.f;
    if (d >= 6) {
      int i = d - 6;
      float tgt = fmaf(act[i], M.ascale[i], M.aoffset[i]);
      float e = qpos[7 + i] + qvel[6 + i] * M.h - tgt;
      t = -M.kp[i] * e - M.kd[i] * qvel[6 + i];
    }
    tin[d] = t;
  }
  __syncwarp();
  aba_inward<false, true, false, false>(M, L, sm, lane, sm + L.spd_ax, sm + L.spd_xpos, sm + L.spd_U, sm + L.spd_Dinv, nullptr, tin);
  aba_outward(M, L, sm, lane, sm + L.spd_ax, sm + L.spd_xpos, sm + L.spd_U, sm + L.spd_Dinv, sm + L.qstar);
  for (int i = lane; i < M.nu; i += 32) {
    float tgt = fmaf(act[i], M.ascale[i], M.aoffset[i]);
    float e = qpos[7 + i] + qvel[6 + i] * M.h - tgt;
    float a = sm[L.spd_ab + 6 + i] + sm[L.qstar + 6 + i];
    float t = -M.kp[i] * e - M.kd[i] * (qvel[6 + i] + a * M.h);
    tau[i] = fminf(fmaxf(t, -M.tlim[i]), M.tlim[i]);
  }
  __syncwarp();
}

// ------------------------------------------------------------------ mj_forward at the current state (tau given) -> qacc, sensors, contact mask
struct FwdOut { unsigned long long mask; int iters; int status; };
__device__ FwdOut forward_dynamics(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  FwdOut o; o.status = 0;
  fk_pass<true>(M, L, sm, lane);
  o.mask = collide(M, L, sm, lane);
  int nlim = make_limits(M, L, sm, lane);
  o.iters = solve_constrained(M, L, sm, lane, (o.mask != 0ull) || (nlim > 0));
  // framelinvel / frameangvel of every body origin (pre-integration, quirk Q2)
  for (int b = lane; b < M.nb; b += 32) {
    S6 v = ld6(sm + L.vel + 6 * b);
    st3(sm + L.sens + 6 * b, v.l + cross(v.a, ld3(sm + L.xpos + 3 * b)));
    st3(sm + L.sens + 6 * b + 3, v.a);
  }
  __syncwarp();
  return o;
}

// semi-implicit Euler (mj_Euler, A.9); lanes 0..2 return the root displacement of this substep
__device__ float integrate(const DevModel& M, const EnvLayout& L, float* sm, int lane) {
  float *qpos = sm + L.qpos, *qvel = sm + L.qvel, *qacc = sm + L.qacc, *qwarm = sm + L.qwarm;
  float h = M.h, disp = 0.f;
  for (int d = lane; d < M.nv; d += 32) {
    float a = qacc[d], v = fmaf(h, a, qvel[d]);
    qvel[d] = v; qwarm[d] = a;
    if (d < 3) { disp = h * v; qpos[d] += disp; }
    else if (d >= 6) qpos[d + 1] = fmaf(h, v, qpos[d + 1]);
  }
  __syncwarp();
  if (lane == 0) {
    V3 w = ld3(qvel + 3);
    float n = sqrtf(dot(w, w)), ang = n * h;
    Q4 q; q.w = qpos[3]; q.x = qpos[4]; q.y = qpos[5]; q.z = qpos[6];
    if (ang > 0.f) {
      float sn, cs; sincosf(0.5f * ang, &sn, &cs);
      float s = sn / n;
      Q4 dq; dq.w = cs; dq.x = w.x * s; dq.y = w.y * s; dq.z = w.z * s;
      q = qmul(q, dq);
    }
    q = qnormalize(q);
    qpos[3] = q.w; qpos[4] = q.x; qpos[5] = q.y; qpos[6] = q.z;
  }
  __syncwarp();
  return disp;
}
