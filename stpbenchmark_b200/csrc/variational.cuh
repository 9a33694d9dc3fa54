// placeholder: variational (SVGP) campaign routines -- filled in below
#pragma once
#include "exact.cuh"
