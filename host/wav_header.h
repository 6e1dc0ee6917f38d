// Compatibility name: the reference includes "wav_header.h".
#pragma once
#include "mavg_wav.h"
