// Stand-in include path <math/affinespace.h> (Embree common/ on the include path); see oracle/ref_shim/README.md.
#pragma once
#include "../common/math/affinespace.h"
