// Shadows <cuda_runtime.h> when the engine's CUDA sources are compiled for the CPU emulator (tests/emu/build_emu.py).
#pragma once
#include "cuda_emu.h"
