// chain_model.cuh -- schedule-ordered model table for the chain-lane kernels (v2).
//
// Mapping: LPE = 4 lanes per env, 8 envs per warp.  The kinematic tree (root split into a
// massless "root-trans" pseudo-body with the 3 translational dofs and a "root-rot" pseudo-body
// with the 3 rotational dofs, so every pseudo-body carries <= 3 dofs) is list-scheduled onto the
// 4 lanes: at inward step t lane c processes pseudo-body sched[t][c] (or idles).  The outward
// sweeps run the same table backwards.  Everything a lane produces for its own bodies lives in
// lane-private local memory indexed by the step t (uniform index across the warp -> coalesced,
// L1/L2 backed); only tree junctions whose two sides sit on different lanes go through small
// shared-memory mailboxes.
#pragma once
#include <stdint.h>

#define CH_LPE 4
#define CH_EPW 8          // envs per warp
#define CH_MAXC 4         // contact slots per pseudo-body (one geom: box 4 | capsule 2 | sphere 1)
#define CH_EDGE_MBOX 1000 // in/out edge ids >= this refer to shared-memory mailboxes

#define CH_KIND_HINGE 0
#define CH_KIND_ROOTROT 1
#define CH_KIND_ROOTTRANS 2
#define CH_KIND_ROOT6 3      // un-split free-joint root (6 dofs) -- thread-per-env kernels

struct ChainEntry {       // one (step, lane) cell; 4-byte words only (staged into shared memory as-is)
  int pb;                 // pseudo-body id, -1 = idle
  int body;               // MuJoCo-order body index (root-trans: 0 as well), for outputs
  int kind, ndof, dofadr; // dofadr: index of the first dof in qvel order
  int par_t;              // slot (step) of the parent if it is on this lane, else -1
  int par_mbox;           // mailbox holding the parent's FK / acceleration data if it is on another lane, else -1
  int out_mbox;           // mailbox to publish this body's FK / acceleration data to (cross-lane children), else -1
  int carry_in, carry_out;
  int in_edge[3];         // incoming articulated-inertia edges (-1 none | local edge id | CH_EDGE_MBOX + mailbox edge id)
  int out_edge;           // where Ia/pa go when not carried (-1: carried or tree root)
  int geom, gtype;        // MuJoCo-order geom index (robot geoms, 0-based) or -1
  int limited;            // bit k: dof k is range-limited
  int par_body;           // MuJoCo-order parent body (-1: tree root)
  float bpos[3], bquat[4], mass, ipos[3], inertia[6], tran_iw0;
  float axis[9], arm[3], diw0[3], range[6];
  float kp[3], kd[3], tlim[3], ascale[3], aoffset[3];
  float gpos[3], gmat[9], gsize[3];
};

struct ChainConsts {      // scalars, passed by value (constant bank)
  int T, nb, nq, nv, nu, ng, n_mbox, n_xedge, mb_stride;   // mb_stride: words of mailbox storage per env
  int obs_dim, self_obs_dim;
  float plane_pos[3], plane_n[3], t1_default[3];
  float margin, mu, impratio, solimp[5], imp_a, imp_b, K, B, h, grav[3];
  unsigned long long legal_mask;
  SmplsimEnvCfg cfg;
};
