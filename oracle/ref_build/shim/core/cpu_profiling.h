// shim for src/core/cpu_profiling.h with ENABLE_CPU_PROFILING=0 (the Physics-Lib configuration, premake5.lua:454): every macro is empty
#pragma once
#define CPU_PROFILE_BLOCK(name)
#define CPU_PRINT_PROFILE_BLOCK(name)
#define CPU_PROFILE_STAT(label, value)
