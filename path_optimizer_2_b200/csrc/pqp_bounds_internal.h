// pqp_bounds_internal.h — what other translation units of the library may ask a pqp_bounds_handle
// (defined in pqp_bounds.cu): the device-resident map view and the device it lives on.
#pragma once
#include "../../include/pqp_bounds.h"
#include "pqp_bounds_core.cuh"

namespace pqb {
const MapView *handle_map(const pqp_bounds_handle *h);
int handle_device(const pqp_bounds_handle *h);
}  // namespace pqb
