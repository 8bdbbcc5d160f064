/* Stub for the un-vendored cpuinfo dependency: the reference only mentions it
 * inside an `#if ... && 0` block (src/core/gpu_sw_rasterizer.cpp:112-118). */
#pragma once
extern "C" bool cpuinfo_has_x86_avx2();
