// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Force-included when compiling the UNMODIFIED reference sources in
// /root/reference/src/model/cpp/*.cpp against torch >= 2.x, whose
// <torch/extension.h> no longer declares torch::linalg::inv (used at
// reference src/model/cpp/string.cpp:175). Nothing else is changed.
#pragma once
#include <torch/extension.h>
namespace torch { namespace linalg {
inline at::Tensor inv(const at::Tensor& a) { return at::linalg_inv(a); }
}}  // namespace torch::linalg
