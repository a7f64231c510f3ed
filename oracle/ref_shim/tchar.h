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
static inline HMODULE LoadLibrary(const char *) { return (HMODULE)0; }
static inline void *GetProcAddress(HMODULE, const char *) { return (void *)0; }
static inline BOOL FreeLibrary(HMODULE) { return TRUE; }

#include "nvEncodeAPI.h"
static inline bool operator==(const GUID &a, const GUID &b) { return memcmp(&a, &b, sizeof(GUID)) == 0; }
static inline bool operator!=(const GUID &a, const GUID &b) { return !(a == b); }
#endif
