// hittable.h — forwarding header: the reference's `#include "hittable.h"` resolves to the
// host-side mirror of its scene API (see rtow_host.h).
#pragma once
#include "rtow_host.h"
