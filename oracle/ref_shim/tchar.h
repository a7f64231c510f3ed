/*
 * TEST INFRASTRUCTURE ONLY -- stand-in for MSVC <tchar.h>, pulled in by the
 * reference's nv_enc/nv_enc.cpp:13 before its own headers.  Supplies the few
 * Win32 loader names and the GUID comparison MSVC gets from <guiddef.h>, so the
 * reference encoder translation unit compiles unmodified under gcc.
 */
#ifndef JMC_ORACLE_TCHAR_SHIM_H
#define JMC_ORACLE_TCHAR_SHIM_H
#include <Windows.h>
#include <dlfcn.h>
#include <string>
using std::string;

#define TEXT(x) x
#define _T(x) x

/* nv_enc/nv_enc.cpp:967,970,979 passes a uint32_t* where the CUDA driver API takes a size_t*
 * (an upstream 64-bit bug MSVC lets through).  Route those three call sites through a helper that
 * does the narrowing, defined in oracle/ref_nvenc_driver.cpp.  The macro is function-like, so the
 * `extern tcuMemAllocPitch *cuMemAllocPitch;` declaration in dynlink_cuda_cuda.h is untouched. */
int jmref_allocpitch_compat(void *dptr, void *pitch_u32, size_t width_bytes, size_t height, unsigned elem);
#define cuMemAllocPitch(a, b, c, d, e) jmref_allocpitch_compat((void *)(a), (void *)(b), (size_t)(c), (size_t)(d), (unsigned)(e))
static inline HMODULE LoadLibrary(const char *) { return (HMODULE)0; }
static inline void *GetProcAddress(HMODULE, const char *) { return (void *)0; }
static inline BOOL FreeLibrary(HMODULE) { return TRUE; }

#include "nvEncodeAPI.h"
static inline bool operator==(const GUID &a, const GUID &b) { return memcmp(&a, &b, sizeof(GUID)) == 0; }
static inline bool operator!=(const GUID &a, const GUID &b) { return !(a == b); }
#endif
