// Stand-in include path <math/vec3.h>; see oracle/ref_shim/README.md.
#pragma once
#include "../common/math/vec.h"
