/* Stub autoconf header needed to compile SPRAL's SSIDS CPU sources in place
 * (test infrastructure only; see oracle/Makefile). */
#pragma once
#define HAVE_STD_ALIGN 1
#define HAVE_SCHED_GETCPU 1
