// TEST INFRASTRUCTURE - forced-include shim used ONLY to compile the reference's own, unmodified
// CUDA op (/root/reference/models/ops/src) against torch 2.11 into oracle/_ref/ (build_ref.py).
//
// The reference calls AT_DISPATCH_FLOATING_TYPES(value.type(), ...) (ms_deform_attn_cuda.cu:64,134);
// torch >= 2.x only accepts an at::ScalarType there.  Re-define the macro so that a
// DeprecatedTypeProperties argument is converted, leaving the reference sources untouched.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace rlipv2_ref_compat {
inline at::ScalarType scalar_type_of(at::ScalarType s) { return s; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace rlipv2_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...)                                   \
  AT_DISPATCH_SWITCH(::rlipv2_ref_compat::scalar_type_of(TYPE), NAME,                \
                     AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
