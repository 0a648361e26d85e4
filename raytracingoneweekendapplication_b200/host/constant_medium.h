// constant_medium.h — forwarding header: the reference's `#include "constant_medium.h"` resolves to the
// host-side mirror of its scene API (see rtow_host.h).
#pragma once
#include "rtow_host.h"
