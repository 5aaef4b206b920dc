// Stand-in for Embree's common/math/affinespace.h (see oracle/ref_shim/README.md).  The reference includes it but uses none of it.
#pragma once
#include "vec.h"
namespace embree {
struct LinearSpace3fa { Vec3fa vx, vy, vz; };
struct AffineSpace3fa { LinearSpace3fa l; Vec3fa p; };
}
