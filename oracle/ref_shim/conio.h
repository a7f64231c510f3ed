/* TEST INFRASTRUCTURE ONLY: stand-in for MSVC <conio.h> (test_nv_dec.cpp:14 polls the keyboard). */
#ifndef JMC_TEST_CONIO_SHIM_H
#define JMC_TEST_CONIO_SHIM_H
static inline int _kbhit(void) { return 0; }
static inline int getch(void) { return 0; }
#endif
