// perlin.h — forwarding header: the reference's `#include "perlin.h"` resolves to the
// host-side mirror of its scene API (see rtow_host.h).
#pragma once
#include "rtow_host.h"
