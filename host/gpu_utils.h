// Compatibility name: the reference includes "gpu_utils.h".
#pragma once
#include "mavg_workspace.h"
