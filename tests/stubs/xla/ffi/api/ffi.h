// TEST STUB -- NOT the XLA FFI headers.
//
// jaxlib (which ships xla/ffi/api/{c_api,api,ffi}.h) is not installable in this image, so
// minppo_b200/csrc/xla_ffi_shim.cc cannot be compiled against the real API here.  This file mocks the SUBSET of the
// public C++ API surface the shim uses -- written from the published XLA FFI documentation, not copied from XLA -- so
// that tests can (1) compile the shim, catching plain C++ errors, and (2) drive its handler bodies (context cache,
// aliasing checks, argument forwarding) on a GPU through tests/stubs/ffi_mock_driver.cc.  It proves nothing about ABI
// compatibility with a real jaxlib; INTEGRATION.md says so.
#pragma once

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

enum DataType { PRED, S32, U32, F32 };
template <DataType dt> struct NativeType;
template <> struct NativeType<PRED> { using type = bool; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<F32> { using type = float; };

template <typename T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  Buffer() = default;
  Buffer(void* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
  Span<int64_t> dimensions() const { return Span<int64_t>{dims_.data(), dims_.size()}; }
  T* typed_data() const { return static_cast<T*>(data_); }
  void* untyped_data() const { return data_; }

 private:
  void* data_ = nullptr;
  std::vector<int64_t> dims_;
};

// Result<Buffer<dt>>: pointer-like access to the pre-allocated result buffer
template <DataType dt>
class ResultBuffer {
 public:
  ResultBuffer() = default;
  explicit ResultBuffer(Buffer<dt> b) : b_(std::move(b)) {}
  const Buffer<dt>* operator->() const { return &b_; }

 private:
  Buffer<dt> b_;
};

template <typename T> struct PlatformStream {};

// Binding builder: the mock only has to make `Ffi::Bind().Ctx<..>().Arg<..>().Attr<T>("name").Ret<..>()` well-formed.
struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// The real macro defines an exported XLA_FFI_Handler symbol; the mock only keeps the implementation referenced.
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" void* name##_mock_symbol() { (void)(binding); return reinterpret_cast<void*>(&impl); }
