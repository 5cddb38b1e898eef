/* stand-in for <curand_kernel.h> in the CPU build of the reference kernels: the types live in cuda_emul.h */
#pragma once
#include "cuda_emul.h"
