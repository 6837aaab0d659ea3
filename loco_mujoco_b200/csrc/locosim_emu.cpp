// Serial CPU emulation of the warp-cooperative engine core (development aid only; see locosim_core.cuh).
// Built by tools/build_emu.sh into scratch/; NOT shipped, NOT loaded by the product path.
#define LS_EMULATE 1
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "locosim_config.h"
#include "locosim_host.h"

struct EmuBase {
  HostModel hm;
  DevModel m;
  SolverOpts so;
  std::vector<int> grf_group; int n_grf = 0;
  virtual ~EmuBase() {}
  virtual void reset(const double* q, const double* v) = 0;
  virtual void step(const double* ctrl, int nsub) = 0;
  virtual void fwd(const double* ctrl) = 0;
  virtual void get(double* q, double* v, double* qacc, double* ws) = 0;
  virtual void set_ws(const double* w) = 0;
  virtual int info(int k) = 0;
  virtual void bind_prm() = 0;
  virtual void contact(int k, double* o) = 0;
  virtual void grf(double* o, int clear) = 0;
  virtual void frames(double* gx, double* bx) = 0;
};
template <class C>
struct EmuT : EmuBase {
  EnvS<C> e;
  EmuT() { memset(&e, 0, sizeof(e)); }
  void reset(const double* q, const double* v) override {
    for (int i = 0; i < m.nv; i++) { e.qpos[i] = (float)q[i]; e.qvel[i] = (float)v[i]; e.qacc_ws[i] = 0; e.qacc[i] = 0; }
  }
  void step(const double* ctrl, int nsub) override {
    for (int i = 0; i < m.nu; i++) e.ctrl[i] = (float)ctrl[i];
    c_models[0] = m;
    physics_substeps(0, e, so, nsub, grf_group.data(), n_grf);
  }
  void fwd(const double* ctrl) override {
    for (int i = 0; i < m.nu; i++) e.ctrl[i] = (float)ctrl[i];
    c_models[0] = m;
    forward(0, e, so);
  }
  void get(double* q, double* v, double* qacc, double* ws) override {
    for (int i = 0; i < m.nv; i++) {
      if (q) q[i] = e.qpos[i];
      if (v) v[i] = e.qvel[i];
      if (qacc) qacc[i] = e.qacc[i];
      if (ws) ws[i] = e.qacc_ws[i];
    }
  }
  void set_ws(const double* w) override { for (int i = 0; i < m.nv; i++) e.qacc_ws[i] = (float)w[i]; }
  int info(int k) override { return k == 0 ? e.ncon : (k == 1 ? e.nefc : e.solver_iter); }
  void contact(int k, double* o) override {
    o[0] = e.con_g1[k]; o[1] = e.con_g2[k]; o[2] = e.con_dim[k]; o[3] = e.con_dist[k];
    int r0 = e.nunit + e.con_row[k], dim = e.con_dim[k];
    for (int j = 0; j < 8; j++) o[4 + j] = (j < (C::CONE == 1 || dim == 1 ? dim : 2 * (dim - 1))) ? e.r_force[r0 + j] : 0.0;
    for (int j = 0; j < 3; j++) { o[12 + j] = e.con_frame[k][j]; o[15 + j] = e.con_pos[k][j]; }
  }
  void grf(double* o, int clear) override { for (int k = 0; k < 3 * LS_MAX_GRF; k++) { o[k] = e.grf[k]; if (clear) e.grf[k] = 0; } }
  void frames(double* gx, double* bx) override {
    for (int g = 0; g < m.ng; g++) for (int k = 0; k < 3; k++) gx[3 * g + k] = e.gxpos[g][k];
    for (int b = 0; b < m.nb; b++) { for (int k = 0; k < 3; k++) bx[7 * b + k] = e.xpos[b][k]; for (int k = 0; k < 4; k++) bx[7 * b + 3 + k] = e.xquat[b][k]; }
  }
  void bind_prm() override { e.prm = hm.default_row.data(); e.pk_tab = m.pair_packed; c_models[0] = m; init_workspace(0, e); }
};

static bool emu_has_boxbox(const HostModel& hm) {
  DevModel v;
  bind_model(v, hm, hm.ints.data(), hm.reals.data());
  for (int p = 0; p < hm.np; p++)
    if (v.geom_type[v.pair_geom[2 * p]] == LS_GEOM_BOX && v.geom_type[v.pair_geom[2 * p + 1]] == LS_GEOM_BOX) return true;
  return false;
}

extern "C" {
EmuBase* emu_create(const int* ints, int n_ints, const double* reals, int n_reals) {
  HostModel hm;
  std::string err = parse_model(hm, ints, n_ints, reals, n_reals);
  if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); return nullptr; }
  EmuBase* s;
  if (hm.cone == 1 && hm.integrator == 0) s = new EmuT<CfgEllEuler>();
  else if (hm.cone == 0 && hm.integrator == 0 && (hm.nv > 18 || emu_has_boxbox(hm))) s = new EmuT<CfgPyrEuler29>();   // (as cfg_fits in locosim.cu)
  else if (hm.cone == 0 && hm.integrator == 0) s = new EmuT<CfgPyrEuler>();
  else if (hm.cone == 0 && hm.integrator == 1) s = new EmuT<CfgPyrRK4>();
  else { fprintf(stderr, "emu: no config\n"); return nullptr; }
  s->hm = hm;
  bind_model(s->m, s->hm, s->hm.ints.data(), s->hm.reals.data());
  s->bind_prm();
  s->so.tolerance = 1e-5f; s->so.ls_tolerance = 0.1f; s->so.max_iter = 20; s->so.ls_iter = 16;
  return s;
}
void emu_destroy(EmuBase* s) { delete s; }
void emu_set_opts(EmuBase* s, float tol, float ls_tol, int max_iter, int ls_iter) {
  s->so.tolerance = tol; s->so.ls_tolerance = ls_tol; s->so.max_iter = max_iter; s->so.ls_iter = ls_iter;
}
void emu_reset(EmuBase* s, const double* qpos, const double* qvel) { s->reset(qpos, qvel); }
void emu_step(EmuBase* s, const double* ctrl, int nsub) { s->step(ctrl, nsub); }
void emu_forward(EmuBase* s, const double* ctrl) { s->fwd(ctrl); }
void emu_get_state(EmuBase* s, double* qpos, double* qvel) { s->get(qpos, qvel, nullptr, nullptr); }
void emu_get_qacc(EmuBase* s, double* qacc) { s->get(nullptr, nullptr, qacc, nullptr); }
void emu_get_ws(EmuBase* s, double* w) { s->get(nullptr, nullptr, nullptr, w); }
void emu_set_ws(EmuBase* s, const double* w) { s->set_ws(w); }
long emu_mpr_stat(int k) { long v[8] = {g_forward_evals, g_mpr_candidates, g_mpr_calls, g_mpr_supports, g_sep_found, g_sep_ok, g_mpr_nohit, 0}; return v[k]; }
long emu_ls_evals() { return g_ls_evals; }
long emu_ls_searches() { return g_ls_searches; }
int emu_ncon(EmuBase* s) { return s->info(0); }
void emu_contact(EmuBase* s, int k, double* o) { s->contact(k, o); }
void emu_set_grf(EmuBase* s, const int* group, int ng, int n_grf) { s->grf_group.assign(group, group + ng); s->n_grf = n_grf; }
void emu_get_grf(EmuBase* s, double* o, int clear) { s->grf(o, clear); }
void emu_frames(EmuBase* s, double* gx, double* bx) { s->frames(gx, bx); }
int emu_nefc(EmuBase* s) { return s->info(1); }
int emu_iter(EmuBase* s) { return s->info(2); }
int emu_sizeof_env(int which) {
  switch (which) {
    case 0: return (int)sizeof(EnvS<CfgEllEuler>);
    case 1: return (int)sizeof(EnvS<CfgPyrEuler>);
    default: return (int)sizeof(EnvS<CfgPyrRK4>);
  }
}
}
