// shim for src/rendering/raytracing.h (D3D12 ray tracing): scene.cpp includes it for the component types it clones
#pragma once
#include "render_stubs.h"
