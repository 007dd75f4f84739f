/* Build shim, see libssh.h in this directory. */
#ifndef ORACLE_SHIM_SFTP_H
#define ORACLE_SHIM_SFTP_H
typedef void *sftp_file;
typedef void *sftp_session;
#endif
