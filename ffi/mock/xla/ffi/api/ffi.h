// MINIMAL STAND-IN for jaxlib's xla/ffi/api/ffi.h — test infrastructure only (tests/test_ffi_shim.py).
// jaxlib (which ships the real header) is not installable in the build image, so ffi/idoc_b200_ffi.cc cannot be compiled
// against it here.  This file declares just the surface the shim uses, with the same names and shapes as the documented
// XLA typed-FFI API (AnyBuffer, Result<>, Error, ErrorOr<>, Span, PlatformStream<>, Ffi::Bind().Ctx/Arg/Ret/Attr,
// XLA_FFI_DEFINE_HANDLER_SYMBOL), and makes the DEFINE macro static_assert that the handler's parameter list matches its
// binding (count, order and types of context / argument / result / attribute slots).  It is NOT used by any product build.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };

template <typename T>
class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_;
  size_t n_;
};

class Error {
 public:
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(std::move(m)); }
  static Error Internal(std::string m) { return Error(std::move(m)); }
  bool success() const { return ok_; }

 private:
  Error() : ok_(true) {}
  explicit Error(std::string m) : ok_(false), msg_(std::move(m)) {}
  bool ok_;
  std::string msg_;
};

class Unexpected {
 public:
  explicit Unexpected(Error e) : e_(std::move(e)) {}
  Error e_;
};

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : has_(true), v_(std::move(v)), e_(Error::Success()) {}  // NOLINT
  ErrorOr(Unexpected u) : has_(false), v_(), e_(std::move(u.e_)) {}     // NOLINT
  bool has_value() const { return has_; }
  bool has_error() const { return !has_; }
  const T& value() const { return v_; }
  const Error& error() const { return e_; }

 private:
  bool has_;
  T v_;
  Error e_;
};

class AnyBuffer {
 public:
  DataType element_type() const { return F32; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(nullptr, 0); }
  void* untyped_data() const { return nullptr; }
  size_t size_bytes() const { return 0; }
  size_t element_count() const { return 0; }
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};

template <typename T>
struct PlatformStream {
  using type = T;
};

template <typename... Ts>
struct Binding {
  template <typename C>
  Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename F>
  static constexpr bool Matches = std::is_invocable_r_v<Error, F, Ts...>;
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                          \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>,                              \
                #sym ": handler parameters do not match Ctx/Arg/Ret/Attr of the binding");         \
  extern "C" void* sym(void*) { return nullptr; }
