// count_tiled.cuh — counting kernel for large n (distance matrix does not fit in shared memory).
#pragma once
#include "common.cuh"
namespace qs {}
