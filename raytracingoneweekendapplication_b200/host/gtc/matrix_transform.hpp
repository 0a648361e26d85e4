// gtc/matrix_transform.hpp — forwarding header for the small glm-compatible subset in rtow_host.h
// (vec2/vec3/vec4/mat4, translate/rotate/scale/radians).
#pragma once
#include "../rtow_host.h"
