// dec_transform.cuh -- stages D2+D3
#pragma once
#include "common.cuh"
