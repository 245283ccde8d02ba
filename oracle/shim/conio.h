/* Stub of the Windows <conio.h> used by /root/reference/src/utils.cpp:4,11-13 so that the
 * reference's serial sources compile on Linux.  Test infrastructure (oracle/_ref build) only. */
#ifndef AZ_SHIM_CONIO_H
#define AZ_SHIM_CONIO_H
static inline int _kbhit(void) { return 0; }
static inline int _getch(void) { return 0; }
/* MSVC's bounded scanf: fscanf_s(f, "%s", buf, n) -> fscanf(f, "%s", buf) */
#define fscanf_s(f, fmt, buf, n) fscanf(f, fmt, buf)
#endif
