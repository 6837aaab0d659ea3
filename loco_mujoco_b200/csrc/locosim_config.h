// Compile-time configurations: one step-kernel instantiation per (friction cone, integrator) family.
// NV / NB / NG / MAXOBS are capacities (model sizes must be <=), MAXCON / MAXROW bound the active contact set.
// CONE: 0 pyramidal, 1 elliptic (mjModel.opt.cone); RK4: 0 Euler, 1 RK4 (mjModel.opt.integrator).
// CONVEX: 1 = the configuration carries the general convex (mjc_Convex / MPR) narrow phase; models with such pairs need it.
#pragma once
struct CfgEllEuler { enum { NV = 18, NB = 14, NG = 40,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 1, RK4 = 0, CONVEX = 1 }; };  // UnitreeA1
struct CfgPyrEuler { enum { NV = 18, NB = 14, NG = 56,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 0, CONVEX = 1 }; };  // Talos
struct CfgPyrRK4   { enum { NV = 19, NB = 12, NG = 100, MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 1, CONVEX = 1 }; };  // Atlas, HumanoidTorque
struct CfgPyrEuler29 { enum { NV = 29, NB = 26, NG = 48,  MAXCON = 16, MAXROW = 64, MAXOBS = 64, CONE = 0, RK4 = 0, CONVEX = 1 }; };  // UnitreeG1 (29 dofs <= 32 lanes)
