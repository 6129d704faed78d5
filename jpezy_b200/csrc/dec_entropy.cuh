// dec_entropy.cuh -- stage D1 (filled in below capi_decode.inc)
#pragma once
#include "common.cuh"
