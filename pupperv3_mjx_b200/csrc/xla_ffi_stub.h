/* xla_ffi_stub.h -- the subset of XLA's typed-FFI C API (xla/ffi/api/c_api.h, jaxlib) that pupper_ffi.cc touches, restated so
 * the handlers compile where jaxlib is absent.  Field order follows the published header (every struct starts with
 * struct_size + extension_start); a build against the real header (-DPUPPER_XLA_FFI_HEADER=...) is the authority. */
#ifndef PUPPER_XLA_FFI_STUB_H_
#define PUPPER_XLA_FFI_STUB_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct XLA_FFI_Extension_Base XLA_FFI_Extension_Base;
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;
typedef struct XLA_FFI_Future XLA_FFI_Future;
typedef struct XLA_FFI_Api XLA_FFI_Api;
typedef enum { XLA_FFI_DataType_INVALID = 0, XLA_FFI_DataType_S32 = 4, XLA_FFI_DataType_U8 = 6, XLA_FFI_DataType_U32 = 8, XLA_FFI_DataType_F32 = 11 } XLA_FFI_DataType;
typedef enum { XLA_FFI_Error_Code_OK = 0, XLA_FFI_Error_Code_INVALID_ARGUMENT = 3, XLA_FFI_Error_Code_INTERNAL = 13 } XLA_FFI_Error_Code;
typedef enum { XLA_FFI_ExecutionStage_INSTANTIATE = 0, XLA_FFI_ExecutionStage_PREPARE = 1, XLA_FFI_ExecutionStage_INITIALIZE = 2, XLA_FFI_ExecutionStage_EXECUTE = 3 } XLA_FFI_ExecutionStage;
typedef struct XLA_FFI_Buffer { size_t struct_size; XLA_FFI_Extension_Base *extension_start; XLA_FFI_DataType dtype; void *data; int64_t rank; int64_t *dims; } XLA_FFI_Buffer;
typedef int32_t XLA_FFI_ArgType;  /* 1 = BUFFER */
typedef int32_t XLA_FFI_RetType;  /* 1 = BUFFER */
typedef int32_t XLA_FFI_AttrType;
typedef struct XLA_FFI_ByteSpan { const char *ptr; size_t len; } XLA_FFI_ByteSpan;
typedef struct XLA_FFI_Args { size_t struct_size; XLA_FFI_Extension_Base *extension_start; int64_t size; XLA_FFI_ArgType *types; void **args; } XLA_FFI_Args;
typedef struct XLA_FFI_Rets { size_t struct_size; XLA_FFI_Extension_Base *extension_start; int64_t size; XLA_FFI_RetType *types; void **rets; } XLA_FFI_Rets;
typedef struct XLA_FFI_Attrs { size_t struct_size; XLA_FFI_Extension_Base *extension_start; int64_t size; XLA_FFI_AttrType *types; XLA_FFI_ByteSpan **names; void **attrs; } XLA_FFI_Attrs;
typedef struct XLA_FFI_CallFrame { size_t struct_size; XLA_FFI_Extension_Base *extension_start; const XLA_FFI_Api *api; XLA_FFI_ExecutionContext *ctx; XLA_FFI_ExecutionStage stage; XLA_FFI_Args args; XLA_FFI_Rets rets; XLA_FFI_Attrs attrs; XLA_FFI_Future *future; } XLA_FFI_CallFrame;
typedef struct XLA_FFI_Error_Create_Args { size_t struct_size; XLA_FFI_Extension_Base *extension_start; const char *message; XLA_FFI_Error_Code errc; } XLA_FFI_Error_Create_Args;
typedef struct XLA_FFI_Stream_Get_Args { size_t struct_size; XLA_FFI_Extension_Base *extension_start; XLA_FFI_ExecutionContext *ctx; void *stream; } XLA_FFI_Stream_Get_Args;
typedef XLA_FFI_Error *XLA_FFI_Error_Create(XLA_FFI_Error_Create_Args *args);
typedef XLA_FFI_Error *XLA_FFI_Stream_Get(XLA_FFI_Stream_Get_Args *args);
typedef struct XLA_FFI_Api_Version { size_t struct_size; XLA_FFI_Extension_Base *extension_start; int major_version; int minor_version; } XLA_FFI_Api_Version;
/* Only the leading members the handlers use; the real struct continues with more function pointers. */
struct XLA_FFI_Api { size_t struct_size; XLA_FFI_Extension_Base *extension_start; XLA_FFI_Api_Version api_version; void *internal_api;
                     XLA_FFI_Error_Create *XLA_FFI_Error_Create; void *XLA_FFI_Error_GetMessage; void *XLA_FFI_Error_Destroy;
                     void *XLA_FFI_Handler_Register; XLA_FFI_Stream_Get *XLA_FFI_Stream_Get; };
#ifdef __cplusplus
}
#endif
#endif
