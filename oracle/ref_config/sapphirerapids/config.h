/* TEST INFRASTRUCTURE.  config.h for compiling the reference's GEMM path as TARGET=SAPPHIRERAPIDS
 * (oracle/build_ref.py).  In the reference this header is generated at build time by getarch;
 * these are the values it emitted for this target in a scratch build (cache geometry and
 * feature macros only -- no reference source code). */
#define OS_LINUX	1
#define ARCH_X86_64	1
#define C_GCC	1
#define __64BIT__	1
#define HAVE_C11	1
#define PTHREAD_CREATE_FUNC	pthread_create
#define BUNDERSCORE	_
#define NEEDBUNDERSCORE	1
#define SAPPHIRERAPIDS
#define L1_CODE_SIZE 32768
#define L1_CODE_ASSOCIATIVE 8
#define L1_CODE_LINESIZE 64
#define L1_DATA_SIZE 49152
#define L1_DATA_ASSOCIATIVE 12
#define L1_DATA_LINESIZE 64
#define L2_SIZE 2097152
#define L2_ASSOCIATIVE 7
#define L2_LINESIZE 64
#define DTB_DEFAULT_ENTRIES 32
#define HAVE_CMOV
#define HAVE_MMX
#define HAVE_SSE
#define HAVE_SSE2
#define HAVE_SSE3
#define HAVE_SSSE3
#define HAVE_SSE4_1
#define HAVE_SSE4_2
#define HAVE_AVX
#define HAVE_AVX2
#define HAVE_AVX512VL
#define HAVE_AVX512BF16
#define HAVE_AMXBF16
#define HAVE_FMA3
#define HAVE_CFLUSH
#define HAVE_HIT 1
#define NUM_SHAREDCACHE 1
#define NUM_CORES 8
#define CORE_SAPPHIRERAPIDS
#define CHAR_CORENAME "SAPPHIRERAPIDS"
#define SLOCAL_BUFFER_SIZE	20480
#define DLOCAL_BUFFER_SIZE	12288
#define CLOCAL_BUFFER_SIZE	12288
#define ZLOCAL_BUFFER_SIZE	8192
#define GEMM_MULTITHREAD_THRESHOLD	4
