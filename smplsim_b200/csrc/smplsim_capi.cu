// smplsim_capi.cu -- extern "C" boundary of libsmplsim_b200.so (see include/smplsim.h).
// Host side: float32 image of the model table, tree schedule, shared-memory layout, launches.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "env_kernels.cuh"
#include "chain_kernels.cuh"
#include "chain_host.hpp"
#include "warp_kernels.cuh"
#include "tpe_kernels.cuh"
#include <cstdlib>
#include <algorithm>

struct SmplsimHandle {
  DevModel hm;        // host copy
  DevModel* dm;       // device copy
  EnvLayout lay;
  int num_envs, device, wpb;
  size_t smem_bytes;
  // v2 (chain-lane kernels)
  bool v2 = false;
  int tmax = 0, wpb2 = 4;
  ChainEntry* d_tab = nullptr;
  ChainConsts kc;
  size_t smem2 = 0;
  std::string v2_why;
  // v3 (level-synchronous, compile-time layout): cls 0 none, 1 SMPL lpe32, 2 SMPL lpe16, 3 SMPL-X, 4 generic
  int v3cls = 0, wpb3 = 4, align3 = 1;
  size_t smem3 = 0;
  // v4 (thread-per-env x chain-per-warp)
  bool v4 = false;
  TpeTable* tab4 = nullptr;     // host copy, uploaded to __constant__ c_tpe when this handle becomes the active one
  float* gs4 = nullptr;
  size_t npad4 = 0, smem4 = 0;
  std::string v4_why;
};

typedef TCfg<24, 75, 12> TC_SMPL;
static const SmplsimHandle* g_tpe_owner = nullptr;
static int tpe_activate(SmplsimHandle* h, cudaStream_t st) {
  if (g_tpe_owner == h) return 0;
  if (cudaMemcpyToSymbolAsync(c_tpe, h->tab4, sizeof(TpeTable), 0, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
  g_tpe_owner = h;
  return 0;
}

typedef WCfg<24, 75, 24, 64, 32> WC_SMPL32;
typedef WCfg<24, 75, 24, 64, 16> WC_SMPL16;
typedef WCfg<24, 75, 24, 64, 8> WC_SMPL8;
typedef WCfg<52, 159, 52, 120, 32> WC_SMPLX;
typedef WCfg<64, 192, 64, 128, 32> WC_GEN;

template <class C>
static int v3_configure(SmplsimHandle* h) {
  constexpr size_t mbytes = ((sizeof(WModel<C>) + 15) / 16) * 16;
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  int best_w = 0, best_wpb = 0, forced = h->wpb3;
  const int cand[] = {16, 14, 12, 8, 7, 6, 4, 3, 2, 1};
  for (int wpb : cand) {
    if (forced > 0 && wpb != forced) continue;
    size_t sm = mbytes + (size_t)wpb * C::EPW * C::total * 4;
    if (sm > (size_t)max_smem) continue;
    if (cudaFuncSetAttribute(k_step3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) continue;
    int nblk = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, k_step3<C>, 32 * wpb, sm) != cudaSuccess) continue;
    if (nblk * wpb > best_w) { best_w = nblk * wpb; best_wpb = wpb; }
  }
  if (!best_wpb) return -1;
  h->wpb3 = best_wpb;
  h->smem3 = mbytes + (size_t)best_wpb * C::EPW * C::total * 4;
  if (cudaFuncSetAttribute(k_step3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem3) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_reset3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem3) != cudaSuccess) return -1;
  return best_w;
}
template <class C>
static void v3_step(SmplsimHandle* h, const WStepArgs& a, cudaStream_t st) {
  int per = h->wpb3 * C::EPW;
  k_step3<C><<<(a.n + per - 1) / per, 32 * h->wpb3, h->smem3, st>>>(h->dm, a);
}
template <class C>
static void v3_reset(SmplsimHandle* h, const WResetArgs& a, cudaStream_t st) {
  int per = h->wpb3 * C::EPW;
  k_reset3<C><<<(a.n + per - 1) / per, 32 * h->wpb3, h->smem3, st>>>(h->dm, a);
}
static void launch_step3(SmplsimHandle* h, const WStepArgs& a, cudaStream_t st) {
  switch (h->v3cls) {
    case 1: v3_step<WC_SMPL32>(h, a, st); break;
    case 2: v3_step<WC_SMPL16>(h, a, st); break;
    case 5: v3_step<WC_SMPL8>(h, a, st); break;
    case 3: v3_step<WC_SMPLX>(h, a, st); break;
    default: v3_step<WC_GEN>(h, a, st); break;
  }
}
static void launch_reset3(SmplsimHandle* h, const WResetArgs& a, cudaStream_t st) {
  switch (h->v3cls) {
    case 1: v3_reset<WC_SMPL32>(h, a, st); break;
    case 2: v3_reset<WC_SMPL16>(h, a, st); break;
    case 5: v3_reset<WC_SMPL8>(h, a, st); break;
    case 3: v3_reset<WC_SMPLX>(h, a, st); break;
    default: v3_reset<WC_GEN>(h, a, st); break;
  }
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                     \
  do {                                                                                                  \
    cudaError_t e_ = (x);                                                                               \
    if (e_ != cudaSuccess) return fail(SMPLSIM_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char* smplsim_last_error(void) { return g_err.c_str(); }
extern "C" int smplsim_version(void) { return 110; }   // 110: SmplsimState += pid_integral / pid_last_error, SmplsimAux += status

static EnvLayout make_layout(const DevModel& m) {
  EnvLayout L;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };   // 16-byte aligned sections
  int nb = m.nb, nv = m.nv, nq = m.nq, nu = m.nu, ns = m.nslot;
  L.qpos = take(nq); L.qvel = take(nv); L.act = take(nu); L.tau = take(nu); L.qacc = take(nv); L.qwarm = take(nv);
  L.xpos = take(3 * nb); L.xquat = take(4 * nb); L.xmat = take(9 * nb); L.ax = take(3 * nv); L.vel = take(6 * nb);
  L.abias = take(6 * nb); L.pb = take(6 * nb); L.irb = take(10 * nb); L.IA = take(21 * nb); L.pA = take(6 * nb);
  L.U = take(6 * nv); L.Dinv = take(nv); L.u = take(nv); L.acc = take(6 * nb);
  L.spd_ax = take(3 * nv); L.spd_xpos = take(3 * nb); L.spd_U = take(6 * nv); L.spd_Dinv = take(nv); L.spd_ab = take(nv);
  L.tin = take(nv); L.dadd = take(nv); L.qstar = take(nv); L.tsk = take(8);
  L.lD = take(nv); L.laref = take(nv); L.lr = take(nv); L.lphi = take(nv); L.lrs = take(nv); L.lflag = take(nv);
  L.cpos = take(3 * ns); L.ct1 = take(3 * ns); L.cD = take(ns); L.caref = take(4 * ns); L.cr = take(4 * ns);
  L.cphi = take(4 * ns); L.crs = take(4 * ns); L.cflag = take(ns);
  L.sens = take(6 * nb);
  L.obs = take(m.obs_dim + 4);
  L.total = o;
  return L;
}

static int obs_dims(const SmplsimModelDesc* s, const SmplsimEnvCfg* c, int* self_dim) {
  int nb = s->nbody;
  int n = (c->root_height_obs ? 1 : 0) + 3 * (nb - 1) + 6 * nb;   // humanoid_env.py:293-299
  n += (c->self_obs_v == 1) ? 3 + 3 + s->nu : 6 * nb;
  *self_dim = n;
  if (c->task == SMPLSIM_TASK_SPEED || c->task == SMPLSIM_TASK_REACH) n += 3;
  if (c->task == SMPLSIM_TASK_GETUP) n += 1;
  return n;
}

extern "C" int smplsim_create(const SmplsimModelDesc* s, const SmplsimEnvCfg* cfg, int num_envs, int cuda_device, SmplsimHandle** out) {
  if (!s || !cfg || !out || num_envs <= 0) return fail(SMPLSIM_EINVAL, "smplsim_create: null argument or num_envs <= 0");
  if (s->nbody > SM_MAXB || s->nv > SM_MAXV || s->ngeom > SM_MAXG || s->nbody < 1)
    return fail(SMPLSIM_EUNSUPPORTED, "model exceeds compiled limits (bodies<=64, dofs<=192, geoms<=64)");
  if (s->nq != s->nv + 1 || s->nu != s->nv - 6 || s->body_dofnum[0] != 6 || s->body_parent[0] != -1)
    return fail(SMPLSIM_EUNSUPPORTED, "model class: one tree rooted at a free joint, hinge joints elsewhere");
  if (cfg->self_obs_v != 1 && cfg->self_obs_v != 2) return fail(SMPLSIM_EINVAL, "self_obs_v must be 1 or 2");
  if (cfg->control_mode < 0 || cfg->control_mode > 3) return fail(SMPLSIM_EINVAL, "control_mode must be uhc_pd|pd|torque|simple_pid");
  if (cfg->task < 0 || cfg->task > 3) return fail(SMPLSIM_EINVAL, "unknown task");
  if (cfg->nsubsteps < 1) return fail(SMPLSIM_EINVAL, "nsubsteps < 1");
  SmplsimHandle* h = new SmplsimHandle();
  DevModel& m = h->hm;
  std::memset(&m, 0, sizeof m);
  m.nb = s->nbody; m.nq = s->nq; m.nv = s->nv; m.nu = s->nu; m.ng = s->ngeom;
  int maxd = 0;
  for (int b = 0; b < m.nb; b++) {
    m.parent[b] = s->body_parent[b]; m.dofadr[b] = s->body_dofadr[b]; m.dofnum[b] = s->body_dofnum[b];
    if (b > 0 && (m.parent[b] < 0 || m.parent[b] >= b)) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "bodies must be listed parent-first"); }
    if (b > 0 && m.dofnum[b] > 3) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "more than 3 hinges on a body"); }
    m.depth[b] = b == 0 ? 0 : m.depth[m.parent[b]] + 1;
    if (m.depth[b] > maxd) maxd = m.depth[b];
    for (int k = 0; k < 3; k++) { m.bpos[b][k] = (float)s->body_pos[3 * b + k]; m.ipos[b][k] = (float)s->body_ipos[3 * b + k]; }
    for (int k = 0; k < 4; k++) m.bquat[b][k] = (float)s->body_quat[4 * b + k];
    for (int k = 0; k < 6; k++) m.inertia[b][k] = (float)s->body_inertia[6 * b + k];
    m.mass[b] = (float)s->body_mass[b];
    m.tran_iw0[b] = (float)s->body_invweight0[2 * b];
    for (int k = 0; k < m.dofnum[b]; k++) m.dof_body[m.dofadr[b] + k] = b;
  }
  m.nlevel = maxd + 1;
  if (m.nlevel > SM_MAXL) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "tree deeper than 16 levels"); }
  { int o = 0; for (int l = 0; l < m.nlevel; l++) { m.level_adr[l] = o; for (int b = 0; b < m.nb; b++) if (m.depth[b] == l) m.level_list[o++] = b; } m.level_adr[m.nlevel] = o; }
  { int o = 0; for (int b = 0; b < m.nb; b++) { m.child_adr[b] = o; for (int c = b + 1; c < m.nb; c++) if (m.parent[c] == b) m.child_list[o++] = c; } m.child_adr[m.nb] = o; }
  for (int d = 0; d < m.nv; d++) {
    for (int k = 0; k < 3; k++) m.axis[d][k] = (float)s->dof_axis[3 * d + k];
    m.arm[d] = (float)s->dof_armature[d]; m.diw0[d] = (float)s->dof_invweight0[d];
    m.range[d][0] = (float)s->dof_range[2 * d]; m.range[d][1] = (float)s->dof_range[2 * d + 1];
    m.limited[d] = s->dof_limited[d];
  }
  int ns = 0;
  m.legal_mask = 1ull;
  for (int g = 0; g < m.ng; g++) {
    m.gtype[g] = s->geom_type[g]; m.gbody[g] = s->geom_body[g];
    int mc = m.gtype[g] == SMPLSIM_GEOM_BOX ? 4 : m.gtype[g] == SMPLSIM_GEOM_CAPSULE ? 2 : m.gtype[g] == SMPLSIM_GEOM_SPHERE ? 1 : -1;
    if (mc < 0 || m.gbody[g] < 0 || m.gbody[g] >= m.nb) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "geom type (box|capsule|sphere) / body"); }
    m.slot_adr[g] = ns;
    for (int k = 0; k < mc; k++) { if (ns >= SM_MAXSLOT) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "too many contact slots"); } m.slot_geom[ns++] = g; }
    for (int k = 0; k < 3; k++) { m.gpos[g][k] = (float)s->geom_pos[3 * g + k]; m.gsize[g][k] = (float)s->geom_size[3 * g + k]; }
    for (int k = 0; k < 9; k++) m.gmat[g][k] = (float)s->geom_mat[9 * g + k];
    if (s->geom_legal[g]) m.legal_mask |= 1ull << (g + 1);
  }
  m.slot_adr[m.ng] = ns; m.nslot = ns;
  { int o = 0; for (int b = 0; b < m.nb; b++) { m.bgeom_adr[b] = o; for (int g = 0; g < m.ng; g++) if (m.gbody[g] == b) m.bgeom_list[o++] = g; } m.bgeom_adr[m.nb] = o; }
  double nn = std::sqrt(s->plane_normal[0] * s->plane_normal[0] + s->plane_normal[1] * s->plane_normal[1] + s->plane_normal[2] * s->plane_normal[2]);
  double n[3] = {s->plane_normal[0] / nn, s->plane_normal[1] / nn, s->plane_normal[2] / nn};
  for (int k = 0; k < 3; k++) { m.plane_pos[k] = (float)s->plane_pos[k]; m.plane_n[k] = (float)n[k]; m.grav[k] = (float)s->gravity[k]; }
  { // mju_makeFrame default tangent for this normal
    double t[3] = {0, 0, 0};
    if (n[1] < 0.5 && n[1] > -0.5) t[1] = 1; else t[2] = 1;
    double d = n[0] * t[0] + n[1] * t[1] + n[2] * t[2];
    for (int k = 0; k < 3; k++) t[k] -= d * n[k];
    double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    for (int k = 0; k < 3; k++) m.t1_default[k] = (float)(t[k] / tn);
  }
  m.margin = (float)s->margin; m.mu = (float)s->friction[0]; m.impratio = (float)s->impratio;
  for (int k = 0; k < 5; k++) m.solimp[k] = (float)s->solimp[k];
  { double mid = s->solimp[3], pw = s->solimp[4];
    m.imp_a = (float)(1.0 / std::pow(mid, pw - 1)); m.imp_b = (float)(1.0 / std::pow(1 - mid, pw - 1)); }
  { double dmax = s->solimp[1], tc = s->solref[0], dr = s->solref[1];
    if (tc <= 0) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "direct solref (negative) is not supported"); }
    if (tc < 2 * s->timestep) tc = 2 * s->timestep;   // refsafe
    m.K = (float)(1.0 / std::fmax(1e-15, dmax * dmax * tc * tc * dr * dr)); m.B = (float)(2.0 / std::fmax(1e-15, dmax * tc)); }
  m.h = (float)s->timestep;
  for (int i = 0; i < m.nu; i++) {
    m.kp[i] = (float)s->act_kp[i]; m.kd[i] = (float)s->act_kd[i]; m.tlim[i] = (float)s->act_torque_lim[i];
    m.ascale[i] = (float)s->act_scale[i]; m.aoffset[i] = (float)s->act_offset[i];
  }
  m.cfg = *cfg;
  if (cfg->task == SMPLSIM_TASK_REACH && (cfg->reach_body < 0 || cfg->reach_body >= m.nb)) { delete h; return fail(SMPLSIM_EINVAL, "reach_body out of range"); }
  m.obs_dim = obs_dims(s, cfg, &m.self_obs_dim);
  { // 4-slot list schedule of the inward sweep: children strictly before parents, deepest bodies first
    std::vector<int> step(m.nb, -1);
    int done = 0, t = 0;
    m.sched_T = 0;
    while (done < m.nb && t < SM_MAXSCHED) {
      std::vector<int> ready;
      for (int b = 0; b < m.nb; b++) {
        if (step[b] >= 0) continue;
        bool ok = true;
        for (int ci = m.child_adr[b]; ci < m.child_adr[b + 1]; ci++) { int c = m.child_list[ci]; if (step[c] < 0 || step[c] >= t) ok = false; }
        if (ok) ready.push_back(b);
      }
      std::sort(ready.begin(), ready.end(), [&](int a, int b2) { return m.depth[a] != m.depth[b2] ? m.depth[a] > m.depth[b2] : a < b2; });
      for (int k = 0; k < 4; k++) m.sched[t][k] = k < (int)ready.size() ? ready[k] : -1;
      for (int k = 0; k < 4 && k < (int)ready.size(); k++) { step[ready[k]] = t; done++; }
      t++;
    }
    if (done == m.nb) m.sched_T = t;
    for (int b = 0; b < m.nb; b++) if (m.bgeom_adr[b + 1] - m.bgeom_adr[b] > 1) m.sched_T = 0;   // row-parallel path: one geom per body
    for (int tt = 0; tt < m.sched_T; tt++) {
      int nd = 0, nc = 0, ns = 0;
      for (int k = 0; k < 4; k++) {
        int b = m.sched[tt][k];
        if (b < 0) continue;
        nd = std::max(nd, m.dofnum[b]); nc = std::max(nc, m.child_adr[b + 1] - m.child_adr[b]);
        if (m.bgeom_adr[b + 1] > m.bgeom_adr[b]) { int g = m.bgeom_list[m.bgeom_adr[b]]; ns = std::max(ns, m.slot_adr[g + 1] - m.slot_adr[g]); }
      }
      m.sched_nd[tt] = nd; m.sched_nc[tt] = nc; m.sched_ns[tt] = ns;
    }
    const char* ws = std::getenv("SMPLSIM_WARMSET");
    m.warmset = ws ? std::atoi(ws) : 1;
    { const char* dp = std::getenv("SMPLSIM_DIRTYPATH"); m.dirtypath = dp ? std::atoi(dp) : 1; }
    { const char* lt = std::getenv("SMPLSIM_LS_TOL"); m.ls_tol = lt ? (float)std::atof(lt) : 1e-6f; }
    const char* rp = std::getenv("SMPLSIM_ROWS");
    m.rowpar = rp ? std::atoi(rp) : 0;   // opt-in: measured slower than the level sweeps in round 1 (profiles/r1_k_step3_rows.md)
  }
  h->lay = make_layout(m);
  h->num_envs = num_envs; h->device = cuda_device;
  h->wpb = SM_WARPS_PER_BLOCK;
  cudaError_t e = cudaSetDevice(cuda_device);
  if (e != cudaSuccess) { delete h; return fail(SMPLSIM_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
  int max_smem = 0;
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cuda_device);
  while (h->wpb > 1 && (size_t)h->wpb * h->lay.total * 4 > (size_t)max_smem) h->wpb >>= 1;
  h->smem_bytes = (size_t)h->wpb * h->lay.total * 4;
  if (h->smem_bytes > (size_t)max_smem) { delete h; return fail(SMPLSIM_EUNSUPPORTED, "per-env scratch exceeds shared memory"); }
  e = cudaMalloc(&h->dm, sizeof(DevModel));
  if (e == cudaSuccess) e = cudaMemcpy(h->dm, &m, sizeof(DevModel), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_reset, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_kinematics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes);
  if (e != cudaSuccess) { if (h->dm) cudaFree(h->dm); delete h; return fail(SMPLSIM_ECUDA, std::string("smplsim_create: ") + cudaGetErrorString(e)); }
  // ---- v2: chain-lane kernels (4 lanes per env) when the model fits their schedule limits
  {
    const char* force = std::getenv("SMPLSIM_KERNEL");
    ChainPlan P = chain_plan(s, CH_MAXLEDGE);
    if (!P.ok) h->v2_why = P.why;
    else if (P.T > 16) h->v2_why = "schedule longer than 16 steps";
    else if (cfg->control_mode == SMPLSIM_CTRL_SIMPLE_PID) h->v2_why = "simple_pid is only built for the v1 / v3 kernels";
    else if (!force || std::string(force) != "v2") h->v2_why = "chain-lane kernels are opt-in (SMPLSIM_KERNEL=v2)";
    else {
      ChainConsts& k = h->kc;
      std::memset(&k, 0, sizeof k);
      k.T = P.T; k.nb = m.nb; k.nq = m.nq; k.nv = m.nv; k.nu = m.nu; k.ng = m.ng; k.n_mbox = P.n_mbox; k.n_xedge = P.n_xedge;
      k.mb_stride = (19 * P.n_mbox + 27 * P.n_xedge + CH_SC_WORDS) | 1;
      k.obs_dim = m.obs_dim; k.self_obs_dim = m.self_obs_dim;
      for (int i = 0; i < 3; i++) { k.plane_pos[i] = m.plane_pos[i]; k.plane_n[i] = m.plane_n[i]; k.t1_default[i] = m.t1_default[i]; k.grav[i] = m.grav[i]; }
      k.margin = m.margin; k.mu = m.mu; k.impratio = m.impratio;
      for (int i = 0; i < 5; i++) k.solimp[i] = m.solimp[i];
      k.imp_a = m.imp_a; k.imp_b = m.imp_b; k.K = m.K; k.B = m.B; k.h = m.h; k.legal_mask = m.legal_mask; k.cfg = m.cfg;
      h->tmax = P.T <= 10 ? 10 : 16;
      size_t tabw = (size_t)P.T * CH_LPE * (sizeof(ChainEntry) / 4);
      h->wpb2 = 4;
      h->smem2 = (tabw + (size_t)h->wpb2 * CH_EPW * k.mb_stride) * 4;
      cudaError_t e2 = cudaMalloc(&h->d_tab, P.tab.size() * sizeof(ChainEntry));
      if (e2 == cudaSuccess) e2 = cudaMemcpy(h->d_tab, P.tab.data(), P.tab.size() * sizeof(ChainEntry), cudaMemcpyHostToDevice);
      if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_step2<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem2);
      if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_reset2<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem2);
      if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_step2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem2);
      if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_reset2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem2);
      if (e2 != cudaSuccess) { cudaFree(h->dm); if (h->d_tab) cudaFree(h->d_tab); delete h; return fail(SMPLSIM_ECUDA, std::string("smplsim_create(v2): ") + cudaGetErrorString(e2)); }
      h->v2 = true;
    }
  }
  // ---- v4: thread-per-env x chain-per-warp (SMPL-sized models)
  {
    const char* force = std::getenv("SMPLSIM_KERNEL");
    bool want = force ? std::string(force) == "v4" : false;
    ChainPlan P = chain_plan(s, 0, false);
    if (!want) h->v4_why = "not selected";
    else if (h->v2) h->v4_why = "v2 forced";
    else if (cfg->control_mode == SMPLSIM_CTRL_SIMPLE_PID) h->v4_why = "simple_pid is only built for the v1 / v3 kernels";
    else if (!P.ok) h->v4_why = P.why;
    else if (P.T > TPE_MAXT) h->v4_why = "schedule longer than 16 steps";
    else if (m.nb != TC_SMPL::NB || m.nv != TC_SMPL::NV) h->v4_why = "model size is not the SMPL class (24 bodies, 75 dofs)";
    else if (P.n_xedge > TC_SMPL::NE) h->v4_why = "too many junction edges";
    else {
      h->tab4 = new TpeTable();
      std::memset(h->tab4, 0, sizeof(TpeTable));
      ChainConsts& k = h->tab4->K;
      k.T = P.T; k.nb = m.nb; k.nq = m.nq; k.nv = m.nv; k.nu = m.nu; k.ng = m.ng; k.n_mbox = 0; k.n_xedge = P.n_xedge; k.mb_stride = 0;
      k.obs_dim = m.obs_dim; k.self_obs_dim = m.self_obs_dim;
      for (int i = 0; i < 3; i++) { k.plane_pos[i] = m.plane_pos[i]; k.plane_n[i] = m.plane_n[i]; k.t1_default[i] = m.t1_default[i]; k.grav[i] = m.grav[i]; }
      k.margin = m.margin; k.mu = m.mu; k.impratio = m.impratio;
      for (int i = 0; i < 5; i++) k.solimp[i] = m.solimp[i];
      k.imp_a = m.imp_a; k.imp_b = m.imp_b; k.K = m.K; k.B = m.B; k.h = m.h; k.legal_mask = m.legal_mask; k.cfg = m.cfg;
      for (int t = 0; t < P.T; t++) for (int w = 0; w < TPE_WARPS; w++) h->tab4->e[t][w] = P.tab[(size_t)t * CH_LPE + w];
      for (int t = P.T; t < TPE_MAXT; t++) for (int w = 0; w < TPE_WARPS; w++) h->tab4->e[t][w].pb = -1;
      h->npad4 = ((size_t)num_envs + 31) & ~(size_t)31;
      h->smem4 = (size_t)TC_SMPL::smem_words * 32 * 4;
      cudaError_t e4 = cudaMalloc(&h->gs4, (size_t)TC_SMPL::g_words * h->npad4 * 4);
      if (e4 == cudaSuccess) e4 = cudaMemset(h->gs4, 0, (size_t)TC_SMPL::g_words * h->npad4 * 4);
      if (e4 == cudaSuccess) e4 = cudaFuncSetAttribute(k_step4<TC_SMPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem4);
      if (e4 == cudaSuccess) e4 = cudaFuncSetAttribute(k_reset4<TC_SMPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem4);
      if (e4 != cudaSuccess) { h->v4_why = std::string("cuda: ") + cudaGetErrorString(e4); cudaGetLastError(); if (h->gs4) cudaFree(h->gs4); h->gs4 = nullptr; delete h->tab4; h->tab4 = nullptr; }
      else h->v4 = true;
    }
  }
  // ---- v3: default hot path
  {
    const char* force = std::getenv("SMPLSIM_KERNEL");
    const char* lpe = std::getenv("SMPLSIM_LPE");
    const char* al = std::getenv("SMPLSIM_ALIGN");
    if (al) h->align3 = std::atoi(al);
    if (const char* ag = std::getenv("SMPLSIM_ALIGN_GROUP")) h->align3 = (h->align3 & 255) | (std::atoi(ag) << 8);   // warps per barrier group
    const char* wp = std::getenv("SMPLSIM_WPB");
    h->wpb3 = wp ? std::atoi(wp) : 0;
    bool want = !force || std::string(force) == "v3" || std::string(force) == "v4";
    if (want && !h->v2 && !h->v4) {
      int cls;
      if (m.nb <= 24 && m.nv <= 75 && m.ng <= 24 && m.nslot <= 64) cls = (lpe && std::string(lpe) == "16") ? 2 : (lpe && std::string(lpe) == "8") ? 5 : 1;
      else if (m.nb <= 52 && m.nv <= 159 && m.ng <= 52 && m.nslot <= 120) cls = 3;
      else cls = 4;
      h->v3cls = cls;
      int r = cls == 1 ? v3_configure<WC_SMPL32>(h) : cls == 2 ? v3_configure<WC_SMPL16>(h) : cls == 5 ? v3_configure<WC_SMPL8>(h) : cls == 3 ? v3_configure<WC_SMPLX>(h) : v3_configure<WC_GEN>(h);
      if (r < 0) h->v3cls = 0;
    }
  }
  *out = h;
  return SMPLSIM_OK;
}

extern "C" int smplsim_destroy(SmplsimHandle* h) {
  if (!h) return SMPLSIM_OK;
  cudaSetDevice(h->device);
  cudaFree(h->dm);
  if (h->d_tab) cudaFree(h->d_tab);
  if (h->gs4) cudaFree(h->gs4);
  if (g_tpe_owner == h) g_tpe_owner = nullptr;
  delete h->tab4;
  delete h;
  return SMPLSIM_OK;
}
extern "C" int smplsim_obs_dim(const SmplsimHandle* h) { return h ? h->hm.obs_dim : SMPLSIM_EINVAL; }
extern "C" int smplsim_num_envs(const SmplsimHandle* h) { return h ? h->num_envs : SMPLSIM_EINVAL; }
extern "C" int smplsim_smem_bytes_per_env(const SmplsimHandle* h) {
  if (!h) return SMPLSIM_EINVAL;
  switch (h->v3cls) { case 1: return WC_SMPL32::total * 4; case 2: return WC_SMPL16::total * 4; case 5: return WC_SMPL8::total * 4; case 3: return WC_SMPLX::total * 4; case 4: return WC_GEN::total * 4; }
  return h->lay.total * 4;
}
extern "C" int smplsim_warps_per_block(const SmplsimHandle* h) { return h ? (h->v3cls ? h->wpb3 : h->v2 ? h->wpb2 : h->wpb) : SMPLSIM_EINVAL; }
/* 2: chain-lane kernels (4 lanes/env), 1: generic warp-per-env kernels; steps of the chain schedule */
extern "C" int smplsim_kernel_version(const SmplsimHandle* h) { return h ? (h->v4 ? 4 : h->v2 ? 2 : h->v3cls ? 3 : 1) : SMPLSIM_EINVAL; }
extern "C" int smplsim_schedule_steps(const SmplsimHandle* h) { return h ? (h->v2 ? h->kc.T : h->hm.nlevel) : SMPLSIM_EINVAL; }

static dim3 grid2(const SmplsimHandle* h, int n) { int per = h->wpb2 * CH_EPW; return dim3((n + per - 1) / per); }
static void launch_step2(SmplsimHandle* h, const ChainStepArgs& a, cudaStream_t st) {
  if (h->tmax == 10) k_step2<10><<<grid2(h, a.n), 32 * h->wpb2, h->smem2, st>>>(h->d_tab, h->kc, a);
  else k_step2<16><<<grid2(h, a.n), 32 * h->wpb2, h->smem2, st>>>(h->d_tab, h->kc, a);
}
static void launch_reset2(SmplsimHandle* h, const ChainResetArgs& a, cudaStream_t st) {
  if (h->tmax == 10) k_reset2<10><<<grid2(h, a.n), 32 * h->wpb2, h->smem2, st>>>(h->d_tab, h->kc, a);
  else k_reset2<16><<<grid2(h, a.n), 32 * h->wpb2, h->smem2, st>>>(h->d_tab, h->kc, a);
}

static bool state_ok(const SmplsimState* st) {
  return st && st->qpos && st->qvel && st->qpos_fwd && st->qvel_fwd && st->qacc_warm && st->task_target && st->task_change_step &&
         st->progress && st->recovery && st->rng_counter;
}
static dim3 grid_for(const SmplsimHandle* h, int n) { return dim3((n + h->wpb - 1) / h->wpb); }

extern "C" int smplsim_step(SmplsimHandle* h, const SmplsimState* st, const float* action_dev, float* obs_dev, float* reward_dev,
                            uint8_t* terminated_dev, uint8_t* truncated_dev, const SmplsimAux* aux, void* stream) {
  if (!h || !state_ok(st) || !action_dev) return fail(SMPLSIM_EINVAL, "smplsim_step: null handle/state/action");
  if (h->hm.cfg.control_mode == SMPLSIM_CTRL_SIMPLE_PID && (!st->pid_integral || !st->pid_last_error))
    return fail(SMPLSIM_EINVAL, "smplsim_step: control_mode simple_pid needs state.pid_integral / pid_last_error");
  StepArgs a; std::memset(&a, 0, sizeof a);
  a.st = *st; if (aux) a.aux = *aux;
  a.action = action_dev; a.obs = obs_dev; a.reward = reward_dev; a.terminated = terminated_dev; a.truncated = truncated_dev;
  a.n = h->num_envs; a.nsub = h->hm.cfg.nsubsteps; a.mode = 0;
  if (h->v4) {
    if (tpe_activate(h, (cudaStream_t)stream)) return fail(SMPLSIM_ECUDA, "constant table upload failed");
    TpeStepArgs b; b.st = a.st; b.aux = a.aux; b.action = a.action; b.obs = a.obs; b.reward = a.reward; b.terminated = a.terminated;
    b.truncated = a.truncated; b.gs = h->gs4; b.npad = h->npad4; b.n = a.n; b.nsub = a.nsub; b.mode = 0;
    k_step4<TC_SMPL><<<(a.n + 31) / 32, 128, h->smem4, (cudaStream_t)stream>>>(b);
  } else if (h->v3cls) {
    WStepArgs b; b.st = a.st; b.aux = a.aux; b.action = a.action; b.obs = a.obs; b.reward = a.reward; b.terminated = a.terminated;
    b.truncated = a.truncated; b.n = a.n; b.nsub = a.nsub; b.mode = 0; b.align = h->align3;
    launch_step3(h, b, (cudaStream_t)stream);
  } else if (h->v2) {
    ChainStepArgs b; b.st = a.st; b.aux = a.aux; b.action = a.action; b.obs = a.obs; b.reward = a.reward; b.terminated = a.terminated;
    b.truncated = a.truncated; b.n = a.n; b.nsub = a.nsub; b.mode = 0;
    launch_step2(h, b, (cudaStream_t)stream);
  } else
  k_step<<<grid_for(h, a.n), 32 * h->wpb, h->smem_bytes, (cudaStream_t)stream>>>(h->dm, h->lay, a);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

extern "C" int smplsim_mj_step(SmplsimHandle* h, const SmplsimState* st, const float* ctrl_dev, int nsub, const SmplsimAux* aux, void* stream) {
  if (!h || !state_ok(st) || !ctrl_dev || nsub < 1) return fail(SMPLSIM_EINVAL, "smplsim_mj_step: bad argument");
  StepArgs a; std::memset(&a, 0, sizeof a);
  a.st = *st; if (aux) a.aux = *aux;
  a.action = ctrl_dev; a.n = h->num_envs; a.nsub = nsub; a.mode = 1;
  if (h->v4) {
    if (tpe_activate(h, (cudaStream_t)stream)) return fail(SMPLSIM_ECUDA, "constant table upload failed");
    TpeStepArgs b; std::memset(&b, 0, sizeof b);
    b.st = a.st; b.aux = a.aux; b.action = a.action; b.gs = h->gs4; b.npad = h->npad4; b.n = a.n; b.nsub = a.nsub; b.mode = 1;
    k_step4<TC_SMPL><<<(a.n + 31) / 32, 128, h->smem4, (cudaStream_t)stream>>>(b);
  } else if (h->v3cls) {
    WStepArgs b; std::memset(&b, 0, sizeof b);
    b.st = a.st; b.aux = a.aux; b.action = a.action; b.n = a.n; b.nsub = a.nsub; b.mode = 1; b.align = h->align3;
    launch_step3(h, b, (cudaStream_t)stream);
  } else if (h->v2) {
    ChainStepArgs b; std::memset(&b, 0, sizeof b);
    b.st = a.st; b.aux = a.aux; b.action = a.action; b.n = a.n; b.nsub = a.nsub; b.mode = 1;
    launch_step2(h, b, (cudaStream_t)stream);
  } else
  k_step<<<grid_for(h, a.n), 32 * h->wpb, h->smem_bytes, (cudaStream_t)stream>>>(h->dm, h->lay, a);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

extern "C" int smplsim_reset(SmplsimHandle* h, const SmplsimState* st, const uint8_t* mask_dev, int init_mode, const float* qpos0_dev,
                             const float* qvel0_dev, float* obs_dev, const SmplsimAux* aux, void* stream) {
  if (!h || !state_ok(st)) return fail(SMPLSIM_EINVAL, "smplsim_reset: null handle/state");
  if (h->hm.cfg.control_mode == SMPLSIM_CTRL_SIMPLE_PID && (!st->pid_integral || !st->pid_last_error))
    return fail(SMPLSIM_EINVAL, "smplsim_reset: control_mode simple_pid needs state.pid_integral / pid_last_error (Fall init runs the controller)");
  int mode = init_mode < 0 ? h->hm.cfg.state_init : init_mode;
  if (mode < 0 || mode > 2) return fail(SMPLSIM_EINVAL, "smplsim_reset: init_mode");
  if (mode == SMPLSIM_INIT_MOCAP && (!qpos0_dev || !qvel0_dev)) return fail(SMPLSIM_EINVAL, "smplsim_reset: MoCap init needs qpos0/qvel0");
  ResetArgs a; std::memset(&a, 0, sizeof a);
  a.st = *st; if (aux) a.aux = *aux;
  a.mask = mask_dev; a.qpos0 = qpos0_dev; a.qvel0 = qvel0_dev; a.obs = obs_dev; a.n = h->num_envs; a.init_mode = mode;
  if (h->v4) {
    if (tpe_activate(h, (cudaStream_t)stream)) return fail(SMPLSIM_ECUDA, "constant table upload failed");
    TpeResetArgs b; b.st = a.st; b.aux = a.aux; b.mask = a.mask; b.qpos0 = a.qpos0; b.qvel0 = a.qvel0; b.obs = a.obs; b.gs = h->gs4; b.npad = h->npad4;
    b.n = a.n; b.init_mode = a.init_mode;
    k_reset4<TC_SMPL><<<(a.n + 31) / 32, 128, h->smem4, (cudaStream_t)stream>>>(b);
  } else if (h->v3cls) {
    WResetArgs b; b.st = a.st; b.aux = a.aux; b.mask = a.mask; b.qpos0 = a.qpos0; b.qvel0 = a.qvel0; b.obs = a.obs; b.n = a.n; b.init_mode = a.init_mode;
    launch_reset3(h, b, (cudaStream_t)stream);
  } else if (h->v2) {
    ChainResetArgs b; b.st = a.st; b.aux = a.aux; b.mask = a.mask; b.qpos0 = a.qpos0; b.qvel0 = a.qvel0; b.obs = a.obs; b.n = a.n; b.init_mode = a.init_mode;
    launch_reset2(h, b, (cudaStream_t)stream);
  } else
  k_reset<<<grid_for(h, a.n), 32 * h->wpb, h->smem_bytes, (cudaStream_t)stream>>>(h->dm, h->lay, a);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

extern "C" int smplsim_kinematics(SmplsimHandle* h, const float* qpos_dev, float* xpos_dev, float* xquat_dev, int n, void* stream) {
  if (!h || !qpos_dev || !xpos_dev || !xquat_dev || n < 0) return fail(SMPLSIM_EINVAL, "smplsim_kinematics: bad argument");
  if (n == 0) return SMPLSIM_OK;
  k_kinematics<<<grid_for(h, n), 32 * h->wpb, h->smem_bytes, (cudaStream_t)stream>>>(h->dm, h->lay, qpos_dev, xpos_dev, xquat_dev, n);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

extern "C" int smplsim_self_obs(SmplsimHandle* h, int version, const float* qvel_dev, const float* xpos_dev, const float* xquat_dev,
                                const float* linvel_dev, const float* angvel_dev, float* obs_dev, int n, void* stream) {
  if (!h || !xpos_dev || !xquat_dev || !obs_dev || n < 0) return fail(SMPLSIM_EINVAL, "smplsim_self_obs: bad argument");
  if (version == 1 && !qvel_dev) return fail(SMPLSIM_EINVAL, "smplsim_self_obs: v1 needs qvel");
  if (version == 2 && (!linvel_dev || !angvel_dev)) return fail(SMPLSIM_EINVAL, "smplsim_self_obs: v2 needs body velocities");
  if (version != 1 && version != 2) return fail(SMPLSIM_EINVAL, "smplsim_self_obs: version");
  if (n == 0) return SMPLSIM_OK;
  int nb = h->hm.nb;
  int self_dim = (h->hm.cfg.root_height_obs ? 1 : 0) + 3 * (nb - 1) + 6 * nb + (version == 1 ? 6 + h->hm.nu : 6 * nb);
  int wpb = 4;
  k_self_obs<<<(n + wpb - 1) / wpb, 32 * wpb, wpb * 16 * SM_MAXB * 4, (cudaStream_t)stream>>>(h->dm, version, qvel_dev, xpos_dev, xquat_dev,
                                                                                            linvel_dev, angvel_dev, obs_dev, n, self_dim);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

extern "C" int smplsim_motion_gather(SmplsimHandle* h, const int32_t* motion_ids_dev, const float* motion_times_dev, int n,
                                     const float* motion_len_dev, const int32_t* num_frames_dev, const float* motion_dt_dev,
                                     const int32_t* length_starts_dev, int num_tables, const float* const* tables_dev,
                                     const int32_t* widths, float* const* outs_dev, int32_t* frame_idx_dev, void* stream) {
  if (!motion_ids_dev || !motion_times_dev || !motion_len_dev || !num_frames_dev || !motion_dt_dev || !length_starts_dev || n < 0)
    return fail(SMPLSIM_EINVAL, "smplsim_motion_gather: null argument");
  if (num_tables < 0 || num_tables > SM_MAXTABLES) return fail(SMPLSIM_EINVAL, "smplsim_motion_gather: at most 16 tables");
  if (n == 0) return SMPLSIM_OK;
  (void)h;
  GatherArgs a; std::memset(&a, 0, sizeof a);
  a.ids = motion_ids_dev; a.times = motion_times_dev; a.mlen = motion_len_dev; a.nframes = num_frames_dev; a.mdt = motion_dt_dev;
  a.starts = length_starts_dev; a.frame_idx = frame_idx_dev; a.n = n; a.ntab = num_tables;
  for (int k = 0; k < num_tables; k++) { a.tables[k] = tables_dev[k]; a.outs[k] = outs_dev[k]; a.widths[k] = widths[k]; }
  int wpb = 8;
  k_motion_gather<<<(n + wpb - 1) / wpb, 32 * wpb, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

// ------------------------------------------------------------------ GAE: learning_utils.estimate_advantages:198-218 as a reverse scan
// per env column of the [T,N] rollout; thread per env, coalesced across envs at every t.
__global__ void k_gae(const float* __restrict__ rew, const float* __restrict__ not_done, const float* __restrict__ not_dead,
                      const float* __restrict__ val, const float* __restrict__ next_val, float gamma, float tau, int T, int N,
                      float* __restrict__ adv, float* __restrict__ ret) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float prev_v = next_val ? next_val[e] : 0.f, prev_a = 0.f;
  for (int t = T - 1; t >= 0; t--) {
    size_t i = (size_t)t * N + e;
    float v = val[i];
    float delta = rew[i] + gamma * prev_v * not_dead[i] - v;
    float a = delta + gamma * tau * prev_a * not_done[i];
    adv[i] = a;
    ret[i] = v + a;
    prev_v = v; prev_a = a;
  }
}

extern "C" int smplsim_gae(const float* rewards_dev, const float* not_done_dev, const float* not_dead_dev, const float* values_dev,
                           const float* next_value_dev, float gamma, float tau, int T, int N, float* adv_dev, float* ret_dev, void* stream) {
  if (!rewards_dev || !not_done_dev || !not_dead_dev || !values_dev || !adv_dev || !ret_dev || T < 0 || N < 0)
    return fail(SMPLSIM_EINVAL, "smplsim_gae: null argument");
  if (T == 0 || N == 0) return SMPLSIM_OK;
  k_gae<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards_dev, not_done_dev, not_dead_dev, values_dev, next_value_dev, gamma, tau, T, N, adv_dev, ret_dev);
  CUDA_TRY(cudaGetLastError());
  return SMPLSIM_OK;
}

#ifdef SMPLSIM_TRACE
extern "C" int smplsim_debug_trace(float* out, int maxn) {
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int));
  if (n > 8192) n = 8192;
  if (n > maxn) n = maxn;
  cudaMemcpyFromSymbol(out, g_trace, sizeof(float) * n);
  int z = 0;
  cudaMemcpyToSymbol(g_trace_n, &z, sizeof(int));
  return n;
}
#endif
