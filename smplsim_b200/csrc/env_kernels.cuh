// env_kernels.cuh -- env-level kernels: step / mj_step / reset + observation pack.
//
//   k_step   BaseEnv.step: pre_physics_step (task update) -> physics_step (nsub x [torque + mj_step])
//            -> post_physics_step (obs, reward, reset flags)      base_env.py:86-100, humanoid_env.py:439-469
//   k_reset  HumanoidTask.reset -> HumanoidEnv.reset -> BaseEnv.reset          humanoid_env.py:471-512
// One env per warp, SM_WARPS_PER_BLOCK envs per CTA, state staged in shared memory for the whole call
// (all nsub substeps run on-chip; HBM sees each state/IO tensor exactly once per env step).
#pragma once
#include "physics.cuh"

struct StepArgs {
  SmplsimState st;
  SmplsimAux aux;
  const float* action;   // [N,nu] action (mode 0) or ctrl torque (mode 1)
  float* obs;            // [N,obs_dim]
  float* reward;         // [N]
  uint8_t* terminated;   // [N]
  uint8_t* truncated;    // [N]
  int n, nsub, mode;     // mode 0: env step, 1: raw mj_step with ctrl
};

struct ResetArgs {
  SmplsimState st;
  SmplsimAux aux;
  const uint8_t* mask;   // [N] or NULL
  const float* qpos0;    // [N,nq] (MoCap)
  const float* qvel0;    // [N,nv]
  float* obs;            // [N,obs_dim] or NULL
  int n, init_mode;
};

// task scalars in shared memory (EnvLayout::tsk)
#define TSK_TARGET 0
#define TSK_CHANGE 4
#define TSK_CURT 5
#define TSK_RECOV 6
#define TSK_RNG 7

__device__ __forceinline__ void load_row(float* dst, const float* src, int n, int lane) {
  for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

__device__ void load_task(const DevModel& M, const EnvLayout& L, float* sm, int lane, const SmplsimState& st, int env) {
  if (lane == 0) {
    float* t = sm + L.tsk;
    int* ti = (int*)t;
    for (int j = 0; j < 4; j++) t[TSK_TARGET + j] = st.task_target[4 * env + j];
    ti[TSK_CHANGE] = st.task_change_step[env];
    ti[TSK_CURT] = st.progress[env];
    ti[TSK_RECOV] = st.recovery[env];
    ti[TSK_RNG] = (int)st.rng_counter[env];
  }
  __syncwarp();
}
__device__ void store_task(const DevModel& M, const EnvLayout& L, float* sm, int lane, const SmplsimState& st, int env) {
  __syncwarp();
  if (lane == 0) {
    float* t = sm + L.tsk;
    int* ti = (int*)t;
    for (int j = 0; j < 4; j++) st.task_target[4 * env + j] = t[TSK_TARGET + j];
    st.task_change_step[env] = ti[TSK_CHANGE];
    st.progress[env] = ti[TSK_CURT];
    st.recovery[env] = ti[TSK_RECOV];
    st.rng_counter[env] = (uint32_t)ti[TSK_RNG];
  }
}

// reset_task() of the three tasks (humanoid_speed.py:97-103, humanoid_reach.py:80-90, humanoid_getup.py:83-89); lane 0 only
__device__ void reset_task(const DevModel& M, const EnvLayout& L, float* sm, int env) {
  const SmplsimEnvCfg& c = M.cfg;
  if (c.task == SMPLSIM_TASK_NONE) return;
  float* t = sm + L.tsk;
  int* ti = (int*)t;
  uint32_t r[4];
  philox4x32((uint32_t)ti[TSK_RNG], (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
  ti[TSK_RNG] = ti[TSK_RNG] + 1;
  if (c.task == SMPLSIM_TASK_SPEED) t[0] = (float)(c.tar_speed_max - c.tar_speed_min) * u01(r[0]) + (float)c.tar_speed_min;
  else if (c.task == SMPLSIM_TASK_REACH) {
    t[0] = (float)c.tar_dist_max * (2.0f * u01(r[0]) - 1.0f);
    t[1] = (float)c.tar_dist_max * (2.0f * u01(r[1]) - 1.0f);
    t[2] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[2]) + (float)c.tar_height_min;
  } else t[0] = (float)(c.tar_height_max - c.tar_height_min) * u01(r[0]) + (float)c.tar_height_min;
  ti[TSK_CHANGE] = ti[TSK_CURT] + rand_range(r[3], c.change_steps_min, c.change_steps_max);
}

// mj_checkPos / mj_checkVel (what 0) and mj_checkAcc (what 1) with mj_resetData, SURVEY A.2 (see w_check in warp_kernels.cuh)
__device__ int check_state(const DevModel& M, const EnvLayout& L, float* sm, int lane, int what) {
  bool b0 = false, b1 = false;
  if (what == 0) {
    for (int i = lane; i < M.nq; i += 32) b0 |= !(fabsf(sm[L.qpos + i]) <= 1e10f);
    for (int i = lane; i < M.nv; i += 32) b1 |= !(fabsf(sm[L.qvel + i]) <= 1e10f);
  } else {
    for (int i = lane; i < M.nv; i += 32) b0 |= !(fabsf(sm[L.qacc + i]) <= 1e10f);
  }
  b0 = __any_sync(0xffffffffu, b0); b1 = __any_sync(0xffffffffu, b1);
  int bits = (what == 0) ? (b0 ? 1 : (b1 ? 2 : 0)) : (b0 ? 4 : 0);
  if (bits) {
    for (int i = lane; i < M.nq; i += 32) sm[L.qpos + i] = (i < 3) ? M.bpos[0][i] : (i < 7) ? M.bquat[0][i - 3] : 0.f;
    for (int i = lane; i < M.nv; i += 32) { sm[L.qvel + i] = 0.f; sm[L.qacc + i] = 0.f; sm[L.qwarm + i] = 0.f; }
    for (int i = lane; i < M.nu; i += 32) sm[L.tau + i] = 0.f;
  }
  __syncwarp();
  return bits;
}

// nsub x [compute_torque + mj_step] on the staged state.  ctrl_mode: 0 = controller on `act`, 1 = tau preset.
// Returns (lanes 0..2) the root displacement; mask/iters of the last forward pass through `last`.
__device__ float run_substeps(const DevModel& M, const EnvLayout& L, float* sm, int lane, int nsub, int raw, FwdOut& last,
                              const SmplsimState& st, int env, bool write_fwd, bool prep_last) {
  const bool spd = (M.cfg.control_mode == SMPLSIM_CTRL_UHC_PD);
  const bool stale = M.cfg.spd_stale != 0;
  float disp = 0.f;
  bool restore = false;
  for (int s = 0; s < nsub; s++) {
    bool did_fk = false;
    if (!raw) {
      if (spd && !stale) { fk_pass<true>(M, L, sm, lane); spd_prepare(M, L, sm, lane); did_fk = true; }
      compute_torque(M, L, sm, lane, st, env);
    } else if (restore) {
      for (int i = lane; i < M.nu; i += 32) sm[L.tau + i] = sm[L.act + i];
      restore = false;
      __syncwarp();
    }
    int bad = check_state(M, L, sm, lane, 0);
    if (bad) did_fk = false;
    if (!did_fk) fk_pass<true>(M, L, sm, lane);
    last.mask = collide(M, L, sm, lane);
    int nlim = make_limits(M, L, sm, lane);
    last.iters = solve_constrained(M, L, sm, lane, (last.mask != 0ull) || (nlim > 0));
    if (check_state(M, L, sm, lane, 1)) {   // mj_checkAcc: forward again on the reset data
      bad |= 4;
      fk_pass<true>(M, L, sm, lane);
      last.mask = collide(M, L, sm, lane);
      nlim = make_limits(M, L, sm, lane);
      last.iters = solve_constrained(M, L, sm, lane, (last.mask != 0ull) || (nlim > 0));
    }
    last.status |= bad;
    if (raw && bad) restore = true;
    if (s == nsub - 1) {
      for (int b = lane; b < M.nb; b += 32) {   // sensors of the last forward pass (quirk Q2)
        S6 v = ld6(sm + L.vel + 6 * b);
        st3(sm + L.sens + 6 * b, v.l + cross(v.a, ld3(sm + L.xpos + 3 * b)));
        st3(sm + L.sens + 6 * b + 3, v.a);
      }
      if (write_fwd) {   // mj_data.qM / qfrc_bias after this call describe this (pre-integration) state
        load_row(st.qpos_fwd + (size_t)env * M.nq, sm + L.qpos, M.nq, lane);
        load_row(st.qvel_fwd + (size_t)env * M.nv, sm + L.qvel, M.nv, lane);
      }
    }
    if (spd && stale && !raw && (s < nsub - 1 || prep_last)) spd_prepare(M, L, sm, lane);   // factors at s_k serve substep k+1
    disp += integrate(M, L, sm, lane);
  }
  return disp;
}

// heading-inverse quaternion (np_transform_utils.py:34-57,140-146)
__device__ __forceinline__ Q4 heading_inv(const DevModel& M, Q4 root) {
  if (!M.cfg.upright_start) { Q4 bc; bc.w = 0.5f; bc.x = -0.5f; bc.y = -0.5f; bc.z = -0.5f; root = qmul(root, bc); }
  V3 rd = qrot_ref(root, v3(1.f, 0.f, 0.f));
  float hd = atan2f(rd.y, rd.x), sn, cs;
  sincosf(-0.5f * hd, &sn, &cs);
  Q4 h; h.w = cs; h.x = 0.f; h.y = 0.f; h.z = sn;
  return qnormalize(h);
}

// compute_humanoid_self_obs_v1/_v2 (humanoid_env.py:565-688) into out[] (any memory); xpos relative to `root`.
__device__ void pack_self_obs(const DevModel& M, int version, int lane, float root_h, const float* xpos_rel, const float* xquat,
                              const float* qvel, const float* sens, int sens_stride_ang, float* out) {
  int nb = M.nb;
  Q4 r0; r0.w = xquat[0]; r0.x = xquat[1]; r0.y = xquat[2]; r0.z = xquat[3];
  Q4 hq = heading_inv(M, r0);
  int o = 0;
  if (M.cfg.root_height_obs) { if (lane == 0) out[0] = root_h; o = 1; }
  int o_rot = o + 3 * (nb - 1), o_vel = o_rot + 6 * nb;
  for (int b = lane; b < nb; b += 32) {
    if (b > 0) st3(out + o + 3 * (b - 1), qrot_ref(hq, ld3(xpos_rel + 3 * b)));
    Q4 q; q.w = xquat[4 * b]; q.x = xquat[4 * b + 1]; q.y = xquat[4 * b + 2]; q.z = xquat[4 * b + 3];
    Q4 lq = qmul(hq, q);
    st3(out + o_rot + 6 * b, qrot_ref(lq, v3(1.f, 0.f, 0.f)));
    st3(out + o_rot + 6 * b + 3, qrot_ref(lq, v3(0.f, 0.f, 1.f)));
    if (version == 2) {
      st3(out + o_vel + 3 * b, qrot_ref(hq, ld3(sens + 6 * b)));
      st3(out + o_vel + 3 * nb + 3 * b, qrot_ref(hq, ld3(sens + 6 * b + sens_stride_ang)));
    }
  }
  if (version == 1) {
    if (lane == 0) st3(out + o_vel, qrot_ref(hq, ld3(qvel)));
    if (lane == 1) st3(out + o_vel + 3, qrot_ref(hq, ld3(qvel + 3)));
    for (int i = lane; i < M.nu; i += 32) out[o_vel + 6 + i] = qvel[6 + i];
  }
}

// compute_observations (humanoid_task.py:41-44): self obs + task obs; staged in smem then streamed out coalesced
__device__ void write_obs(const DevModel& M, const EnvLayout& L, float* sm, int lane, float* obs_row) {
  float* ob = sm + L.obs;
  const float* qpos = sm + L.qpos;
  pack_self_obs(M, M.cfg.self_obs_v, lane, qpos[2], sm + L.xpos, sm + L.xquat, sm + L.qvel, sm + L.sens, 3, ob);
  if (lane == 0) {
    const float* t = sm + L.tsk;
    int o = M.self_obs_dim;
    Q4 r0; r0.w = qpos[3]; r0.x = qpos[4]; r0.y = qpos[5]; r0.z = qpos[6];
    if (M.cfg.task == SMPLSIM_TASK_SPEED) {
      V3 d = qrot_ref(heading_inv(M, r0), v3(1.f, 0.f, 0.f));
      ob[o] = d.x; ob[o + 1] = d.y; ob[o + 2] = t[0];
    } else if (M.cfg.task == SMPLSIM_TASK_REACH) {
      st3(ob + o, qrot_ref(heading_inv(M, r0), ld3(t) - ld3(qpos)));
    } else if (M.cfg.task == SMPLSIM_TASK_GETUP) ob[o] = t[0];
  }
  __syncwarp();
  if (obs_row) for (int i = lane; i < M.obs_dim; i += 32) obs_row[i] = ob[i];
}

__device__ void write_aux(const DevModel& M, const EnvLayout& L, float* sm, int lane, const SmplsimAux& aux, int env, const FwdOut& fo) {
  V3 root = ld3(sm + L.qpos);
  if (aux.xpos) for (int b = lane; b < M.nb; b += 32) st3(aux.xpos + ((size_t)env * M.nb + b) * 3, ld3(sm + L.xpos + 3 * b) + root);
  if (aux.xquat) load_row(aux.xquat + (size_t)env * M.nb * 4, sm + L.xquat, 4 * M.nb, lane);
  if (aux.body_linvel) for (int b = lane; b < M.nb; b += 32) st3(aux.body_linvel + ((size_t)env * M.nb + b) * 3, ld3(sm + L.sens + 6 * b));
  if (aux.body_angvel) for (int b = lane; b < M.nb; b += 32) st3(aux.body_angvel + ((size_t)env * M.nb + b) * 3, ld3(sm + L.sens + 6 * b + 3));
  if (aux.qacc) load_row(aux.qacc + (size_t)env * M.nv, sm + L.qacc, M.nv, lane);
  if (aux.ctrl) load_row(aux.ctrl + (size_t)env * M.nu, sm + L.tau, M.nu, lane);
  if (lane == 0) {
    if (aux.contact_mask) aux.contact_mask[env] = fo.mask;
    if (aux.solver_iter) aux.solver_iter[env] = fo.iters;
    if (aux.status) aux.status[env] = (uint8_t)fo.status;
  }
}

// stage SPD factors at the state of the last forward pass (qpos_fwd, qvel_fwd): mj_data.qM / qfrc_bias as the reference's
// controller sees them at the first substep of a step (quirk Q1)
__device__ void stage_spd_from_fwd(const DevModel& M, const EnvLayout& L, float* sm, int lane, const SmplsimState& st, int env) {
  load_row(sm + L.qpos, st.qpos_fwd + (size_t)env * M.nq, M.nq, lane);
  load_row(sm + L.qvel, st.qvel_fwd + (size_t)env * M.nv, M.nv, lane);
  __syncwarp();
  fk_pass<true>(M, L, sm, lane);
  spd_prepare(M, L, sm, lane);
}

extern __shared__ float smem_dyn[];

__global__ void __launch_bounds__(32 * SM_WARPS_PER_BLOCK) k_step(const DevModel* __restrict__ Mp, EnvLayout L, StepArgs a) {
  const DevModel& M = *Mp;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= a.n) return;
  float* sm = smem_dyn + (size_t)wib * L.total;
  const bool spd = (M.cfg.control_mode == SMPLSIM_CTRL_UHC_PD);
  if (spd && M.cfg.spd_stale && a.mode == 0) stage_spd_from_fwd(M, L, sm, lane, a.st, env);
  load_row(sm + L.qpos, a.st.qpos + (size_t)env * M.nq, M.nq, lane);
  load_row(sm + L.qvel, a.st.qvel + (size_t)env * M.nv, M.nv, lane);
  load_row(sm + L.qwarm, a.st.qacc_warm + (size_t)env * M.nv, M.nv, lane);
  load_row(sm + L.act, a.action + (size_t)env * M.nu, M.nu, lane);
  if (a.mode != 0) load_row(sm + L.tau, a.action + (size_t)env * M.nu, M.nu, lane);
  load_task(M, L, sm, lane, a.st, env);
  if (a.mode == 0 && lane == 0) {   // pre_physics_step -> update_task (humanoid_task.py:26-28)
    int* ti = (int*)(sm + L.tsk);
    if (M.cfg.task != SMPLSIM_TASK_NONE && ti[TSK_CURT] >= ti[TSK_CHANGE]) reset_task(M, L, sm, env);
  }
  __syncwarp();
  FwdOut fo; fo.mask = 0ull; fo.iters = 0; fo.status = 0;
  float disp = run_substeps(M, L, sm, lane, a.nsub, a.mode, fo, a.st, env, true, false);
  // ---- post_physics_step
  fk_pass<false>(M, L, sm, lane);   // mj_kinematics at the integrated state (humanoid_env.py:389)
  if (a.mode == 0) {
    int* ti = (int*)(sm + L.tsk);
    if (lane == 0) ti[TSK_CURT] += 1;
    __syncwarp();
    write_obs(M, L, sm, lane, a.obs ? a.obs + (size_t)env * M.obs_dim : nullptr);
    float dx = __shfl_sync(FULLMASK, disp, 0), dy = __shfl_sync(FULLMASK, disp, 1);
    if (lane == 0) {
      const SmplsimEnvCfg& c = M.cfg;
      const float* t = sm + L.tsk;
      float rew = 0.f;
      if (c.task == SMPLSIM_TASK_SPEED) {
        float inv_dt = 1.0f / (M.h * (float)a.nsub);
        float vx = dx * inv_dt, vy = dy * inv_dt, e = t[0] - vx;
        rew = expf(-0.25f * (e * e + 0.1f * vy * vy));
      } else if (c.task == SMPLSIM_TASK_REACH) {
        V3 dlt = ld3(t) - (ld3(sm + L.xpos + 3 * c.reach_body) + ld3(sm + L.qpos));
        rew = expf(-4.0f * dot(dlt, dlt));
      } else if (c.task == SMPLSIM_TASK_GETUP) {
        float e = t[0] - sm[L.qpos + 2];
        rew = expf(-4.0f * e * e);
      }
      int term = 0, trunc = 0, pass_time = ti[TSK_CURT] > c.episode_length;
      if (c.task == SMPLSIM_TASK_NONE) trunc = pass_time;
      else if (c.task == SMPLSIM_TASK_GETUP && ti[TSK_RECOV] > 0) ti[TSK_RECOV] -= 1;
      else { trunc = pass_time; term = (fo.mask & ~M.legal_mask) != 0ull; }
      if (a.reward) a.reward[env] = rew;
      if (a.terminated) a.terminated[env] = (uint8_t)term;
      if (a.truncated) a.truncated[env] = (uint8_t)trunc;
    }
  }
  write_aux(M, L, sm, lane, a.aux, env, fo);
  load_row(a.st.qpos + (size_t)env * M.nq, sm + L.qpos, M.nq, lane);
  load_row(a.st.qvel + (size_t)env * M.nv, sm + L.qvel, M.nv, lane);
  load_row(a.st.qacc_warm + (size_t)env * M.nv, sm + L.qwarm, M.nv, lane);
  if (a.mode == 0) store_task(M, L, sm, lane, a.st, env);
}

__global__ void __launch_bounds__(32 * SM_WARPS_PER_BLOCK) k_reset(const DevModel* __restrict__ Mp, EnvLayout L, ResetArgs a) {
  const DevModel& M = *Mp;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= a.n) return;
  if (a.mask && !a.mask[env]) return;
  float* sm = smem_dyn + (size_t)wib * L.total;
  const SmplsimEnvCfg& c = M.cfg;
  int init = a.init_mode < 0 ? c.state_init : a.init_mode;
  load_task(M, L, sm, lane, a.st, env);
  if (lane == 0) {
    int* ti = (int*)(sm + L.tsk);
    if (c.task == SMPLSIM_TASK_GETUP) ti[TSK_RECOV] = c.recovery_steps;
    if (!c.legacy_change_step) ti[TSK_CURT] = 0;
    reset_task(M, L, sm, env);   // sees the old cur_t when legacy_change_step (quirk Q4)
  }
  for (int i = lane; i < M.nq; i += 32) sm[L.qpos + i] = 0.f;
  for (int i = lane; i < M.nv; i += 32) { sm[L.qvel + i] = 0.f; sm[L.qwarm + i] = 0.f; sm[L.qacc + i] = 0.f; }
  for (int i = lane; i < M.nu; i += 32) sm[L.tau + i] = 0.f;
  __syncwarp();
  FwdOut fo; fo.mask = 0ull; fo.iters = 0; fo.status = 0;
  if (init == SMPLSIM_INIT_DEFAULT) {
    if (lane == 0) { sm[L.qpos + 2] = 0.94f; sm[L.qpos + 3] = 0.5f; sm[L.qpos + 4] = 0.5f; sm[L.qpos + 5] = 0.5f; sm[L.qpos + 6] = 0.5f; }
  } else if (init == SMPLSIM_INIT_FALL) {
    if (lane == 0) { sm[L.qpos + 2] = 0.3f; sm[L.qpos + 3] = 1.0f; }
    __syncwarp();
    if (c.control_mode == SMPLSIM_CTRL_UHC_PD && c.spd_stale) { fk_pass<true>(M, L, sm, lane); spd_prepare(M, L, sm, lane); }  // mj_forward
    int ngrp = (M.nu + 3) / 4;
    for (int k = 0; k < 3; k++) {
      int* ti = (int*)(sm + L.tsk);
      uint32_t base = (uint32_t)ti[TSK_RNG];
      for (int gidx = lane; gidx < ngrp; gidx += 32) {
        uint32_t r[4];
        philox4x32(base + (uint32_t)gidx, (uint32_t)env, 0u, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32), r);
        for (int j = 0; j < 4 && 4 * gidx + j < M.nu; j++) sm[L.act + 4 * gidx + j] = u01(r[j]) - 0.5f;
      }
      __syncwarp();
      if (lane == 0) ti[TSK_RNG] = (int)(base + (uint32_t)ngrp);
      __syncwarp();
      run_substeps(M, L, sm, lane, c.nsubsteps, 0, fo, a.st, env, false, true);
    }
  } else {
    load_row(sm + L.qpos, a.qpos0 + (size_t)env * M.nq, M.nq, lane);
    load_row(sm + L.qvel, a.qvel0 + (size_t)env * M.nv, M.nv, lane);
  }
  __syncwarp();
  // reset_sim(): mj_forward at the reset state -> fresh sensors / contacts; qM, qfrc_bias fresh (qpos_fwd = qpos)
  fk_pass<true>(M, L, sm, lane);
  fo.mask = collide(M, L, sm, lane);
  for (int b = lane; b < M.nb; b += 32) {
    S6 v = ld6(sm + L.vel + 6 * b);
    st3(sm + L.sens + 6 * b, v.l + cross(v.a, ld3(sm + L.xpos + 3 * b)));
    st3(sm + L.sens + 6 * b + 3, v.a);
  }
  if (lane == 0) ((int*)(sm + L.tsk))[TSK_CURT] = 0;
  __syncwarp();
  write_obs(M, L, sm, lane, a.obs ? a.obs + (size_t)env * M.obs_dim : nullptr);
  write_aux(M, L, sm, lane, a.aux, env, fo);
  load_row(a.st.qpos + (size_t)env * M.nq, sm + L.qpos, M.nq, lane);
  load_row(a.st.qvel + (size_t)env * M.nv, sm + L.qvel, M.nv, lane);
  load_row(a.st.qpos_fwd + (size_t)env * M.nq, sm + L.qpos, M.nq, lane);
  load_row(a.st.qvel_fwd + (size_t)env * M.nv, sm + L.qvel, M.nv, lane);
  load_row(a.st.qacc_warm + (size_t)env * M.nv, sm + L.qwarm, M.nv, lane);
  store_task(M, L, sm, lane, a.st, env);
}

// ------------------------------------------------------------------ standalone utility kernels (one env per warp)
// mj_kinematics / poselib global_transformation: qpos -> xpos (world), xquat
__global__ void k_kinematics(const DevModel* __restrict__ Mp, EnvLayout L, const float* qpos, float* xpos, float* xquat, int n) {
  const DevModel& M = *Mp;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= n) return;
  float* sm = smem_dyn + (size_t)wib * L.total;
  load_row(sm + L.qpos, qpos + (size_t)env * M.nq, M.nq, lane);
  __syncwarp();
  fk_pass<false>(M, L, sm, lane);
  V3 root = ld3(sm + L.qpos);
  for (int b = lane; b < M.nb; b += 32) st3(xpos + ((size_t)env * M.nb + b) * 3, ld3(sm + L.xpos + 3 * b) + root);
  load_row(xquat + (size_t)env * M.nb * 4, sm + L.xquat, 4 * M.nb, lane);
}

// compute_humanoid_self_obs_v1/_v2 on caller-supplied body states
__global__ void k_self_obs(const DevModel* __restrict__ Mp, int version, const float* qvel, const float* xpos, const float* xquat,
                           const float* linvel, const float* angvel, float* obs, int n, int self_dim) {
  const DevModel& M = *Mp;
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= n) return;
  float* sm = smem_dyn + (size_t)wib * (16 * SM_MAXB);
  float *xr = sm, *sens = sm + 3 * SM_MAXB;
  const float* xp = xpos + (size_t)env * M.nb * 3;
  V3 root = ld3(xp);
  for (int b = lane; b < M.nb; b += 32) {
    st3(xr + 3 * b, ld3(xp + 3 * b) - root);
    if (version == 2) {
      st3(sens + 6 * b, ld3(linvel + ((size_t)env * M.nb + b) * 3));
      st3(sens + 6 * b + 3, ld3(angvel + ((size_t)env * M.nb + b) * 3));
    }
  }
  __syncwarp();
  pack_self_obs(M, version, lane, root.z, xr, xquat + (size_t)env * M.nb * 4, qvel ? qvel + (size_t)env * M.nv : nullptr, sens, 3,
                obs + (size_t)env * self_dim);
}

// get_motion_state_intervaled: frame = floor(clip(t,0,len)/dt) clipped to the clip, + length_starts[id]; row gather of every table
#define SM_MAXTABLES 16
struct GatherArgs {
  const int32_t* ids; const float* times; const float* mlen; const int32_t* nframes; const float* mdt; const int32_t* starts;
  const float* tables[SM_MAXTABLES]; float* outs[SM_MAXTABLES]; int widths[SM_MAXTABLES];
  int32_t* frame_idx; int n, ntab;
};
__global__ void k_motion_gather(GatherArgs a) {
  int lane = threadIdx.x & 31;
  int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.n) return;
  int id = a.ids[i];
  float len = a.mlen[id], dt = a.mdt[id], t = a.times[i];
  int nf = a.nframes[id];
  // motion_lib_base.py:448-458 + :321-323 (quirk Q12): phase clip, float idx0, blend, int() truncation
  float phase = fminf(fmaxf(t / len, 0.f), 1.f);
  if (t < 0.f) t = 0.f;
  float idx0 = phase * (float)(nf - 1);
  float idx1 = fminf(idx0 + 1.f, (float)(nf - 1));
  float blend = fminf(fmaxf((t - idx0 * dt) / dt, 0.f), 1.f);
  int fr = (int)((1.0f - blend) * idx0 + blend * idx1);
  size_t row = (size_t)(fr + a.starts[id]);
  if (lane == 0 && a.frame_idx) a.frame_idx[i] = (int)row;
  for (int k = 0; k < a.ntab; k++) {
    const float* src = a.tables[k] + row * a.widths[k];
    float* dst = a.outs[k] + (size_t)i * a.widths[k];
    for (int j = lane; j < a.widths[k]; j += 32) dst[j] = src[j];
  }
}
