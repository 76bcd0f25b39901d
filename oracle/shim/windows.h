/* Headless stand-in for <windows.h>: the reference includes it unconditionally on
 * non-Apple builds (raytracing.cpp:10, mesh.cpp:8) and relies on it for memcpy.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#pragma once
#include <string.h>
#include <stdlib.h>
