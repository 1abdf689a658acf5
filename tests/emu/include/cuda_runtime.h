// TEST INFRASTRUCTURE ONLY.  Stands in for <cuda_runtime.h> when the product's .cu sources are compiled for the
// HOST by tests/emu (see cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
