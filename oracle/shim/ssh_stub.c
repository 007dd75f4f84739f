/* Build shim for the oracle: stands in for src/sfio_ssh_session.c (libssh
 * upload/download threads, referenced from src/main.c:298,320,344). None of
 * these is reachable from `-c`/`-x`. */
#include <stddef.h>
void *upload(void *p) { (void)p; return NULL; }
void *download(void *p) { (void)p; return NULL; }
void *remote_decompression(void *p) { (void)p; return NULL; }
