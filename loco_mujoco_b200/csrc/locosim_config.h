// Compile-time configurations: one step-kernel instantiation per (friction cone, integrator) family.
// NV / NB / NG / MAXOBS are capacities (model sizes must be <=), MAXCON / MAXROW bound the active contact set.
// CONE: 0 pyramidal, 1 elliptic (mjModel.opt.cone); RK4: 0 Euler, 1 RK4 (mjModel.opt.integrator).
// CONVEX: 1 = the configuration carries the general convex (mjc_Convex / MPR) narrow phase; models with such pairs need it.
// BOXBOX: 1 = it also carries the exact edge-edge branch of mjc_BoxBox; models with box-box pairs need it (A1, Talos, H1 have none:
//         their kernels stay free of that code - an A/B on one box showed -1 % for A1 from its mere presence).
#pragma once
struct CfgEllEuler { enum { NV = 18, NB = 14, NG = 40,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 1, RK4 = 0, CONVEX = 1, BOXBOX = 0 }; };  // UnitreeA1
struct CfgPyrEuler { enum { NV = 18, NB = 14, NG = 56,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 0, CONVEX = 1, BOXBOX = 0 }; };  // Talos, UnitreeH1
struct CfgPyrRK4   { enum { NV = 19, NB = 12, NG = 100, MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 1, CONVEX = 1, BOXBOX = 1 }; };  // Atlas, HumanoidTorque
struct CfgPyrEuler29 { enum { NV = 29, NB = 26, NG = 48,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 0, CONVEX = 1, BOXBOX = 1 }; };  // UnitreeG1 (29 dofs <= 32 lanes)
