// Stand-in include path <math/math.h> (Embree common/ on the include path); see oracle/ref_shim/README.md.
#pragma once
#include "../common/math/math.h"
