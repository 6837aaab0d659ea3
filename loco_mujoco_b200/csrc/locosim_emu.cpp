// Serial CPU emulation of the warp-cooperative engine core (development aid only; see locosim_core.cuh).
// Built by tools/build_emu.sh into scratch/; NOT shipped, NOT loaded by the product path.
#define LS_EMULATE 1
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "locosim_config.h"
#include "locosim_host.h"

typedef CfgHumanoid EmuCfg;   // the largest capacity config

struct Emu {
  HostModel hm;
  DevModel m;
  EnvS<EmuCfg> e;
  SolverOpts so;
};

extern "C" {
Emu* emu_create(const int* ints, int n_ints, const double* reals, int n_reals) {
  Emu* s = new Emu();
  std::string err = parse_model(s->hm, ints, n_ints, reals, n_reals);
  if (!err.empty()) { fprintf(stderr, "emu: %s\n", err.c_str()); delete s; return nullptr; }
  bind_model(s->m, s->hm, s->hm.ints.data(), s->hm.reals.data());
  memset(&s->e, 0, sizeof(s->e));
  s->so.tolerance = 1e-6f; s->so.ls_tolerance = 0.01f; s->so.max_iter = 8; s->so.ls_iter = 16;
  return s;
}
void emu_destroy(Emu* s) { delete s; }
void emu_set_opts(Emu* s, float tol, float ls_tol, int max_iter, int ls_iter) {
  s->so.tolerance = tol; s->so.ls_tolerance = ls_tol; s->so.max_iter = max_iter; s->so.ls_iter = ls_iter;
}
void emu_reset(Emu* s, const double* qpos, const double* qvel) {
  for (int i = 0; i < s->m.nv; i++) {
    s->e.qpos[i] = (float)qpos[i]; s->e.qvel[i] = (float)qvel[i]; s->e.qacc_ws[i] = 0; s->e.qacc[i] = 0;
  }
}
void emu_step(Emu* s, const double* ctrl, int nsub) {
  for (int i = 0; i < s->m.nu; i++) s->e.ctrl[i] = (float)ctrl[i];
  physics_substeps(s->m, s->e, s->so, nsub);
}
void emu_get_state(Emu* s, double* qpos, double* qvel) {
  for (int i = 0; i < s->m.nv; i++) { qpos[i] = s->e.qpos[i]; qvel[i] = s->e.qvel[i]; }
}
int emu_ncon(Emu* s) { return s->e.ncon; }
int emu_nefc(Emu* s) { return s->e.nefc; }
int emu_iter(Emu* s) { return s->e.solver_iter; }
int emu_sizeof_env(int which) {
  switch (which) {
    case 0: return (int)sizeof(EnvS<CfgA1>);
    case 1: return (int)sizeof(EnvS<CfgAtlas>);
    case 2: return (int)sizeof(EnvS<CfgTalos>);
    default: return (int)sizeof(EnvS<CfgHumanoid>);
  }
}
}
extern "C" {
void emu_forward(Emu* s, const double* ctrl) {
  for (int i = 0; i < s->m.nu; i++) s->e.ctrl[i] = (float)ctrl[i];
  forward(s->m, s->e, s->so);
}
void emu_get_qacc(Emu* s, double* qacc) { for (int i = 0; i < s->m.nv; i++) qacc[i] = s->e.qacc[i]; }
void emu_get_ws(Emu* s, double* w) { for (int i = 0; i < s->m.nv; i++) w[i] = s->e.qacc_ws[i]; }
void emu_set_ws(Emu* s, const double* w) { for (int i = 0; i < s->m.nv; i++) s->e.qacc_ws[i] = (float)w[i]; }
}
