// Compatibility name: the reference includes "benchmark.h".
#pragma once
#include "mavg_bench.h"
