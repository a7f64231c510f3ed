/*
 * TEST INFRASTRUCTURE ONLY -- stand-in for <Windows.h>.
 *
 * The reference (mojing1999/jmcodec) is MSVC-only.  This header gives gcc just
 * enough Win32 vocabulary to compile the reference translation units that hold
 * the CPU surface-format loops, in place, from /root/reference, without editing
 * them (see oracle/Makefile, target _ref/libjmref.so).  Nothing here is product
 * code; none of the threading primitives do anything because the checker only
 * calls the single-threaded conversion functions.
 */
#ifndef JMC_ORACLE_WINDOWS_SHIM_H
#define JMC_ORACLE_WINDOWS_SHIM_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef void *HANDLE;
typedef void *HINSTANCE;
typedef void *HMODULE;
typedef void *LPVOID;
typedef unsigned long DWORD;
typedef int BOOL;
typedef const char *LPCSTR;
typedef DWORD (*LPTHREAD_START_ROUTINE)(LPVOID);

#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define INFINITE 0xFFFFFFFFu
#define WINAPI
#define WAIT_OBJECT_0 0

#define _declspec(x) __attribute__((visibility("default")))
#define __declspec(x) __attribute__((visibility("default")))
#define __stdcall
#define sprintf_s snprintf

static inline HANDLE CreateMutexA(void *, BOOL, LPCSTR) { return (HANDLE)1; }
static inline HANDLE CreateMutex(void *, BOOL, LPCSTR) { return (HANDLE)1; }
static inline HANDLE CreateEventA(void *, BOOL, BOOL, LPCSTR) { return (HANDLE)1; }
static inline HANDLE CreateEvent(void *, BOOL, BOOL, LPCSTR) { return (HANDLE)1; }
static inline DWORD WaitForSingleObject(HANDLE, DWORD) { return WAIT_OBJECT_0; }
static inline BOOL ReleaseMutex(HANDLE) { return TRUE; }
static inline BOOL SetEvent(HANDLE) { return TRUE; }
static inline BOOL ResetEvent(HANDLE) { return TRUE; }
static inline BOOL CloseHandle(HANDLE) { return TRUE; }
static inline HANDLE CreateThread(void *, size_t, LPTHREAD_START_ROUTINE, LPVOID, DWORD, DWORD *) { return (HANDLE)0; }
static inline void Sleep(DWORD ms) { usleep(ms * 1000); }

#endif
