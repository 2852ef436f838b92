/*
 * znzlib.h -- stdio-backed stand-in for nifticlib's znzlib (TEST INFRASTRUCTURE ONLY).
 *
 * The reference vendors nifti1_io.c but not znzlib (it is pulled by CMake's
 * ExternalProject from gitlab.com/slckr/nifticlib @ e26a94e9, see the reference's
 * CMakeLists.txt:115-127).  nifti1_io.c only needs the eight entry points below;
 * this header declares them with the public znzlib names so the reference source
 * compiles unmodified when the oracle build (oracle/Makefile) puts this directory
 * on the include path.  Uncompressed files only (.nii / .hdr+.img).
 */
#ifndef S3D_ORACLE_ZNZLIB_SHIM_H
#define S3D_ORACLE_ZNZLIB_SHIM_H
#include <stdio.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

struct znzptr {
    int withz;
    FILE *nzfptr;
};
typedef struct znzptr *znzFile;

#define znz_isnull(f) ((f) == NULL)
#define znzclose(f) Xznzclose(&(f))

znzFile znzopen(const char *path, const char *mode, int use_compression);
int Xznzclose(znzFile *file);
size_t znzread(void *buf, size_t size, size_t nmemb, znzFile file);
size_t znzwrite(const void *buf, size_t size, size_t nmemb, znzFile file);
long znzseek(znzFile file, long offset, int whence);
int znzrewind(znzFile stream);
long znztell(znzFile file);
int znzputs(const char *str, znzFile file);

#ifdef __cplusplus
}
#endif
#endif
