// chain_host.hpp -- host-side list scheduler and table builder for the chain-lane kernels (v2).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "chain_model.cuh"

struct ChainPlan {
  bool ok = false;
  std::string why;
  int T = 0, n_mbox = 0, n_xedge = 0, n_ledge = 0;
  std::vector<ChainEntry> tab;   // [T][CH_LPE]
};

static ChainPlan chain_plan(const SmplsimModelDesc* s, int max_ledge, bool split_root = true) {
  ChainPlan P;
  const int nb = s->nbody, off = split_root ? 1 : 0, npb = nb + off;
  std::vector<int> par(npb), depth(npb, 0);
  par[0] = -1;
  if (split_root) { par[1] = 0; depth[1] = 1; }
  for (int b = 1; b < nb; b++) { par[b + off] = s->body_parent[b] + off; depth[b + off] = depth[par[b + off]] + 1; }
  std::vector<std::vector<int>> kids(npb);
  for (int p = 1; p < npb; p++) kids[par[p]].push_back(p);
  // one geom per body
  std::vector<int> geom_of(nb, -1);
  for (int g = 0; g < s->ngeom; g++) {
    int b = s->geom_body[g];
    if (geom_of[b] >= 0) { P.why = "more than one geom on a body"; return P; }
    geom_of[b] = g;
  }
  for (int b = 1; b < nb; b++) if (s->body_dofnum[b] > 3) { P.why = "more than 3 hinges on a body"; return P; }
  // ---- list scheduling of the inward sweep (children before parents), CH_LPE lanes
  std::vector<int> step(npb, -1), lane(npb, -1), prev(CH_LPE, -1);
  int scheduled = 0, t = 0;
  while (scheduled < npb) {
    std::vector<int> ready;
    for (int p = 0; p < npb; p++) {
      if (step[p] >= 0) continue;
      bool ok = true;
      for (int k : kids[p]) if (step[k] < 0 || step[k] >= t) ok = false;
      if (ok) ready.push_back(p);
    }
    std::sort(ready.begin(), ready.end(), [&](int a, int b) { return depth[a] != depth[b] ? depth[a] > depth[b] : a < b; });
    if ((int)ready.size() > CH_LPE) ready.resize(CH_LPE);
    std::vector<int> cur(CH_LPE, -1);
    std::vector<char> placed(ready.size(), 0);
    for (size_t i = 0; i < ready.size(); i++)            // carries: continue up the chain on the same lane
      for (int c = 0; c < CH_LPE; c++)
        if (prev[c] >= 0 && par[prev[c]] == ready[i] && cur[c] < 0) { cur[c] = ready[i]; placed[i] = 1; break; }
    for (size_t i = 0; i < ready.size(); i++) {           // affinity: a lane that holds one of the children
      if (placed[i]) continue;
      for (int k : kids[ready[i]]) { int c = lane[k]; if (cur[c] < 0) { cur[c] = ready[i]; placed[i] = 1; break; } }
    }
    for (size_t i = 0; i < ready.size(); i++) {           // anything free (prefer lanes that idled last step)
      if (placed[i]) continue;
      int best = -1;
      for (int c = 0; c < CH_LPE; c++) if (cur[c] < 0 && (best < 0 || (prev[c] < 0 && prev[best] >= 0))) best = c;
      cur[best] = ready[i]; placed[i] = 1;
    }
    for (int c = 0; c < CH_LPE; c++) if (cur[c] >= 0) { step[cur[c]] = t; lane[cur[c]] = c; scheduled++; }
    prev = cur;
    t++;
    if (t > 4096) { P.why = "scheduler did not terminate"; return P; }
  }
  P.T = t;
  // ---- edges and mailboxes
  std::vector<int> out_edge(npb, -1), carry_out(npb, 0), carry_in(npb, 0), out_mbox(npb, -1), par_mbox(npb, -1), par_t(npb, -1);
  std::vector<std::vector<int>> in_edges(npb);
  std::vector<int> nle(CH_LPE, 0);
  for (int p = 1; p < npb; p++) {
    int q = par[p];
    if (lane[p] == lane[q]) {
      par_t[p] = step[q];
      if (step[q] == step[p] + 1 && !carry_in[q]) { carry_out[p] = 1; carry_in[q] = 1; continue; }
      if (!split_root) {   // thread-per-env kernels: every non-carried edge is a shared-memory stash slot
        int id = CH_EDGE_MBOX + P.n_xedge++;
        out_edge[p] = id; in_edges[q].push_back(id);
        continue;
      }
      int id = nle[lane[p]]++;
      out_edge[p] = id; in_edges[q].push_back(id);
    } else {
      if (out_mbox[q] < 0) out_mbox[q] = P.n_mbox++;
      par_mbox[p] = out_mbox[q];
      int id = CH_EDGE_MBOX + P.n_xedge++;
      out_edge[p] = id; in_edges[q].push_back(id);
    }
  }
  for (int c = 0; c < CH_LPE; c++) P.n_ledge = std::max(P.n_ledge, nle[c]);
  if (P.n_ledge > max_ledge) { P.why = "too many same-lane non-adjacent tree edges"; return P; }
  for (int p = 0; p < npb; p++) if (in_edges[p].size() > 3) { P.why = "more than 3 non-carried children"; return P; }
  // ---- table
  P.tab.assign((size_t)P.T * CH_LPE, ChainEntry());
  for (auto& e : P.tab) { std::memset(&e, 0, sizeof e); e.pb = -1; e.par_t = e.par_mbox = e.out_mbox = -1; e.in_edge[0] = e.in_edge[1] = e.in_edge[2] = -1; e.out_edge = -1; e.geom = -1; }
  for (int p = 0; p < npb; p++) {
    ChainEntry& e = P.tab[(size_t)step[p] * CH_LPE + lane[p]];
    e.pb = p;
    int b;
    if (split_root) {
      e.kind = p == 0 ? CH_KIND_ROOTTRANS : p == 1 ? CH_KIND_ROOTROT : CH_KIND_HINGE;
      b = p == 0 ? 0 : p - 1;
      e.ndof = p <= 1 ? 3 : s->body_dofnum[b];
      e.dofadr = p == 0 ? 0 : p == 1 ? 3 : s->body_dofadr[b];
    } else {
      e.kind = p == 0 ? CH_KIND_ROOT6 : CH_KIND_HINGE;
      b = p;
      e.ndof = s->body_dofnum[b];
      e.dofadr = s->body_dofadr[b];
    }
    e.body = b;
    e.par_body = b == 0 ? -1 : s->body_parent[b];
    e.par_t = par_t[p]; e.par_mbox = par_mbox[p]; e.out_mbox = out_mbox[p];
    e.carry_in = carry_in[p]; e.carry_out = carry_out[p]; e.out_edge = out_edge[p];
    for (size_t i = 0; i < in_edges[p].size(); i++) e.in_edge[i] = in_edges[p][i];
    if (p == 0 && split_root) continue;
    for (int k = 0; k < 3; k++) { e.bpos[k] = (float)s->body_pos[3 * b + k]; e.ipos[k] = (float)s->body_ipos[3 * b + k]; }
    for (int k = 0; k < 4; k++) e.bquat[k] = (float)s->body_quat[4 * b + k];
    for (int k = 0; k < 6; k++) e.inertia[k] = (float)s->body_inertia[6 * b + k];
    e.mass = (float)s->body_mass[b];
    e.tran_iw0 = (float)s->body_invweight0[2 * b];
    for (int k = 0; k < e.ndof && k < 3 && e.kind != CH_KIND_ROOT6; k++) {
      int d = e.dofadr + k;
      for (int j = 0; j < 3; j++) e.axis[3 * k + j] = (float)s->dof_axis[3 * d + j];
      e.arm[k] = (float)s->dof_armature[d]; e.diw0[k] = (float)s->dof_invweight0[d];
      e.range[2 * k] = (float)s->dof_range[2 * d]; e.range[2 * k + 1] = (float)s->dof_range[2 * d + 1];
      if (e.kind == CH_KIND_HINGE) {
        if (s->dof_limited[d]) e.limited |= 1 << k;
        int i = d - 6;
        e.kp[k] = (float)s->act_kp[i]; e.kd[k] = (float)s->act_kd[i]; e.tlim[k] = (float)s->act_torque_lim[i];
        e.ascale[k] = (float)s->act_scale[i]; e.aoffset[k] = (float)s->act_offset[i];
      }
    }
    int g = geom_of[b];
    if (g >= 0) {
      e.geom = g; e.gtype = s->geom_type[g];
      if (e.gtype != SMPLSIM_GEOM_BOX && e.gtype != SMPLSIM_GEOM_CAPSULE && e.gtype != SMPLSIM_GEOM_SPHERE) { P.why = "geom type"; return P; }
      for (int k = 0; k < 3; k++) { e.gpos[k] = (float)s->geom_pos[3 * g + k]; e.gsize[k] = (float)s->geom_size[3 * g + k]; }
      for (int k = 0; k < 9; k++) e.gmat[k] = (float)s->geom_mat[9 * g + k];
    }
  }
  P.ok = true;
  return P;
}
