// TEST STUB of the XLA FFI C++ surface (xla/ffi/api/ffi.h of jaxlib) -- just enough of Buffer / ResultBuffer / Error /
// Ffi::Bind() for `g++ -fsyntax-only` to type-check kbot-joystick_b200/csrc/kbs_xla_ffi.cc in an image without jaxlib
// (tests/test_host_cpu.py::test_xla_ffi_shim_compiles_against_stub_headers).  Unlike a bare syntax check, the stub's
// Bind().To(impl) static_asserts that the handler is invocable with exactly the context / attribute / argument / result
// types the binding lists, in order -- the mistake a real build would report first.  Not part of the product.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <utility>

namespace xla { namespace ffi {

enum DataType { F32, U8, S32, S64 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<S64> { using type = int64_t; };

struct Span { const int64_t* p; size_t n; int64_t operator[](size_t i) const { return p[i]; } size_t size() const { return n; } };

template <DataType dtype>
struct Buffer {
  using T = typename NativeType<dtype>::type;
  T* data_ = nullptr; const int64_t* dims_ = nullptr; size_t rank_ = 0; size_t bytes_ = 0;
  T* typed_data() const { return data_; }
  Span dimensions() const { return Span{dims_, rank_}; }
  size_t size_bytes() const { return bytes_; }
};
template <DataType dtype>
struct ResultBuffer {
  Buffer<dtype> b_;
  Buffer<dtype>* operator->() { return &b_; }
  const Buffer<dtype>* operator->() const { return &b_; }
};

struct Error {
  static Error Success() { return Error{}; }
  static Error Internal(const char*) { return Error{}; }
};

template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename C::stream_type_> CtxImpl() const { return {}; }
  template <typename C> auto Ctx() const { return CtxOf<C>::apply(*this); }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> auto Ret() const { return RetOf<T>::apply(*this); }
  template <typename F> int To(F&&) const {
    static_assert(std::is_invocable_r_v<Error, F, Ts...>, "XLA-FFI handler signature does not match its binding");
    return 0;
  }
  template <typename C> struct CtxOf;
  template <typename S> struct CtxOf<PlatformStream<S>> { static Binding<Ts..., S> apply(const Binding&) { return {}; } };
  template <typename T> struct RetOf;
  template <DataType d> struct RetOf<Buffer<d>> { static Binding<Ts..., ResultBuffer<d>> apply(const Binding&) { return {}; } };
};

struct Ffi { static Binding<> Bind() { return {}; } };

}}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding) \
  extern "C" int sym() { static int h = (binding).To(impl); return h; }
