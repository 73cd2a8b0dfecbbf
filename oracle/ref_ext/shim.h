// Build shim for the reference's CUDA sources (test infrastructure only, see oracle/__init__.py).
// The reference calls AT_DISPATCH_FLOATING_TYPES(x.type(), ...); torch >= 2.x resolves the dispatch type through
// ::detail::scalar_type(), which no longer has an overload for at::DeprecatedTypeProperties (what .type() returns).
// Supplying that overload lets the reference's .cu files compile UNMODIFIED, from where they lie under /root/reference.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
