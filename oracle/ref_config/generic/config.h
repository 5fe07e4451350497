/* TEST INFRASTRUCTURE.  config.h for compiling the reference's GEMM path as TARGET=GENERIC
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
#define GENERIC
#define L1_DATA_SIZE 32768
#define L1_DATA_LINESIZE 128
#define L2_SIZE 512488
#define L2_LINESIZE 128
#define DTB_DEFAULT_ENTRIES 128
#define DTB_SIZE 4096
#define L2_ASSOCIATIVE 8
#define CORE_generic
#define CHAR_CORENAME "generic"
#define SLOCAL_BUFFER_SIZE	4096
#define DLOCAL_BUFFER_SIZE	4096
#define CLOCAL_BUFFER_SIZE	8192
#define ZLOCAL_BUFFER_SIZE	8192
#define GEMM_MULTITHREAD_THRESHOLD	4
