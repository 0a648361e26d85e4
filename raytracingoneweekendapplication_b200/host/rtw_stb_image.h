// rtw_stb_image.h — forwarding header: the reference's `#include "rtw_stb_image.h"` resolves to the
// host-side mirror of its scene API (see rtow_host.h).
#pragma once
#include "rtow_host.h"
