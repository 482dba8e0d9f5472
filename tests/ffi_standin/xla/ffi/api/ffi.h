// TEST INFRASTRUCTURE, not a product header: a stand-in for the subset of XLA's
// "xla/ffi/api/ffi.h" (shipped with jax, `jax.ffi.include_dir()`; JAX is not installable in this
// image) that ffi/jrb_xla_ffi.cc uses, with the same names and shapes of the public API, so that
// g++ -fsyntax-only type-checks every handler of that file against include/jrystal_b200.h
// (tests/test_ffi_shim.py).  It checks what a compiler can check without XLA: argument counts and
// types of every jrb_* call, and that each handler's parameter list matches its binding
// (context, attributes, operands, results in order).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
extern "C" cudaError_t cudaMemcpyAsync(void*, const void*, size_t, cudaMemcpyKind, cudaStream_t);

namespace xla {
namespace ffi {

enum DataType { F64, C128 };
template <DataType>
struct NativeOf;
template <>
struct NativeOf<F64> { using type = double; };
template <>
struct NativeOf<C128> { using type = std::complex<double>; };

class Error {
 public:
  static Error Success() { return Error(); }
  static Error Internal(std::string) { return Error(); }
  static Error InvalidArgument(std::string) { return Error(); }
};

template <DataType T>
class Buffer {
 public:
  using native = typename NativeOf<T>::type;
  native* typed_data() const { return nullptr; }
  std::vector<int64_t> dimensions() const { return {}; }
  size_t element_count() const { return 0; }
};

template <DataType T>
class ResultBufferImpl {
 public:
  Buffer<T>* operator->() { return &b_; }
 private:
  Buffer<T> b_;
};
template <DataType T>
using ResultBuffer = ResultBufferImpl<T>;

template <typename T>
struct PlatformStream {};

// Binding: records the C++ parameter types the handler must have, in order
template <typename... Ps>
struct Binding {
  template <typename T>
  constexpr auto Ctx() const { return Binding<Ps..., typename CtxParam<T>::type>(); }
  template <typename T>
  constexpr auto Attr(const char*) const { return Binding<Ps..., T>(); }
  template <typename T>
  constexpr auto Arg() const { return Binding<Ps..., T>(); }
  template <typename T>
  constexpr auto Ret() const { return Binding<Ps..., typename RetParam<T>::type>(); }

  template <typename T>
  struct CtxParam;
  template <typename S>
  struct CtxParam<PlatformStream<S>> { using type = S; };
  template <typename T>
  struct RetParam;
  template <DataType D>
  struct RetParam<Buffer<D>> { using type = ResultBuffer<D>; };
};

struct Ffi {
  static constexpr Binding<> Bind() { return Binding<>(); }
};

template <typename... Ps>
constexpr bool handler_matches(Error (*)(Ps...), Binding<Ps...>) { return true; }

}  // namespace ffi
}  // namespace xla

// the real macro defines an extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*); here: the symbol
// plus a compile-time check that the handler's signature is exactly what the binding describes
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                         \
  static_assert(::xla::ffi::handler_matches(&fn, binding), #symbol);               \
  extern "C" void* symbol(void*) { return nullptr; }
