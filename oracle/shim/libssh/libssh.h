/* Build shim for the oracle: the reference headers include <libssh/libssh.h>
 * (include/Arithmetic_stream.h:26, include/sam_block.h:25) only for its ssh
 * transport, which is off the coding path. Opaque handles are enough. */
#ifndef ORACLE_SHIM_LIBSSH_H
#define ORACLE_SHIM_LIBSSH_H
typedef void *ssh_session;
#endif
