// Compile-time capacity configurations (one kernel instantiation per robot family).
// NV/NB must match the compiled model exactly; NG, MAXCON, MAXROW, MAXOBS are capacities.
#pragma once
struct CfgA1       { enum { NV = 18, NB = 14, NG = 40,  MAXCON = 16, MAXROW = 64, MAXOBS = 40 }; };   // UnitreeA1 (+dir_arrow body folded)
struct CfgAtlas    { enum { NV = 16, NB = 28, NG = 72,  MAXCON = 16, MAXROW = 64, MAXOBS = 40 }; };
struct CfgTalos    { enum { NV = 18, NB = 36, NG = 96,  MAXCON = 16, MAXROW = 64, MAXOBS = 40 }; };
struct CfgHumanoid { enum { NV = 19, NB = 40, NG = 128, MAXCON = 16, MAXROW = 64, MAXOBS = 40 }; };
