// STAND-IN for XLA's header-only FFI API (xla/ffi/api/ffi.h), TEST INFRASTRUCTURE ONLY.
//
// jax / jaxlib are not installable in the image this repository is developed in, so
// jaxoplanet_b200/csrc/jxp_xla_ffi.cc cannot be compiled against the real header there.  This file
// declares the names that translation unit uses -- same namespaces, class and macro names, member
// functions and call signatures as the public XLA FFI API -- with no behaviour behind them, so that
// tests/test_xla_ffi.py can at least type-check every handler against its binding:
// Binding<...>::To(fn) static_asserts that `fn` is callable with exactly the decoded context,
// argument, result and attribute types, in binding order, and returns ffi::Error.
// It is written from the documented API surface, not copied from XLA.  When `jax.ffi.include_dir()`
// exists, the real header is used instead (jaxoplanet_b200/jax_bindings.py::build_ffi).
#pragma once

#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <type_traits>
#include <utility>

extern "C" {
struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;
}

namespace xla::ffi {

enum class DataType : uint8_t { INVALID = 0, PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;

namespace internal {
template <DataType>
struct NativeTypeOf;
template <>
struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <>
struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <>
struct NativeTypeOf<DataType::F32> { using type = float; };
template <>
struct NativeTypeOf<DataType::F64> { using type = double; };
}  // namespace internal
template <DataType dtype>
using NativeType = typename internal::NativeTypeOf<dtype>::type;

template <typename T>
class Span {
 public:
  constexpr Span(T* data, size_t size) : data_(data), size_(size) {}
  constexpr size_t size() const { return size_; }
  constexpr T& operator[](size_t i) const { return data_[i]; }
  constexpr T* begin() const { return data_; }
  constexpr T* end() const { return data_ + size_; }

 private:
  T* data_;
  size_t size_;
};

inline constexpr size_t kDynamicRank = static_cast<size_t>(-1);

template <DataType dtype, size_t rank = kDynamicRank>
class Buffer {
 public:
  NativeType<dtype>* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  Span<const int64_t> dimensions() const { return Span<const int64_t>(dims_, rank_); }
  size_t element_count() const {
    size_t n = 1;
    for (size_t i = 0; i < rank_; ++i) n *= static_cast<size_t>(dims_[i]);
    return n;
  }

 private:
  NativeType<dtype>* data_ = nullptr;
  const int64_t* dims_ = nullptr;
  size_t rank_ = 0;
};

template <typename T>
class Result {
 public:
  T& operator*() { return value_; }
  T* operator->() { return &value_; }

 private:
  T value_;
};
template <DataType dtype, size_t rank = kDynamicRank>
using ResultBuffer = Result<Buffer<dtype, rank>>;

enum class ErrorCode : uint8_t {
  kOk = 0, kCancelled, kUnknown, kInvalidArgument, kDeadlineExceeded, kNotFound, kAlreadyExists, kPermissionDenied,
  kResourceExhausted, kFailedPrecondition, kAborted, kOutOfRange, kUnimplemented, kInternal, kUnavailable, kDataLoss,
  kUnauthenticated
};

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

class ScratchAllocator {
 public:
  ScratchAllocator() = default;
  ScratchAllocator(ScratchAllocator&&) = default;
  ScratchAllocator& operator=(ScratchAllocator&&) = default;
  ScratchAllocator(const ScratchAllocator&) = delete;
  std::optional<void*> Allocate(size_t size, size_t alignment = 1) {
    (void)size;
    (void)alignment;
    return std::nullopt;
  }
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T>
struct CtxDecoded { using type = T; };
template <typename T>
struct CtxDecoded<PlatformStream<T>> { using type = T; };
}  // namespace internal

class HandlerBase {
 public:
  virtual ~HandlerBase() = default;
  XLA_FFI_Error* Call(XLA_FFI_CallFrame*) const { return nullptr; }
};
// keeps the implementation referenced (as the real handler does), so that the C entry points it calls
// show up as undefined symbols of the handler library
template <typename Fn>
class Handler : public HandlerBase {
 public:
  explicit Handler(Fn fn) : fn_(std::move(fn)) {}
  const Fn& fn() const { return fn_; }

 private:
  Fn fn_;
};

template <typename... Ts>
class Binding {
 public:
  template <typename T>
  Binding<Ts..., typename internal::CtxDecoded<T>::type> Ctx() && { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() && { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() && { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(std::string) && { return {}; }
  template <typename Fn>
  HandlerBase* To(Fn&& fn) && {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>,
                  "handler signature does not match its binding (context, arguments, results, attributes in order)");
    return new Handler<std::decay_t<Fn>>(std::forward<Fn>(fn));
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                   \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame* call_frame) {             \
    static ::xla::ffi::HandlerBase* handler = (binding).To(impl);              \
    return handler->Call(call_frame);                                          \
  }
