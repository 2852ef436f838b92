/* stdio implementation of the znzlib subset declared in shim/znzlib.h (test infrastructure). */
#include "znzlib.h"
#include <stdlib.h>
#include <string.h>

znzFile znzopen(const char *path, const char *mode, int use_compression)
{
    (void)use_compression;
    znzFile f = (znzFile)calloc(1, sizeof(struct znzptr));
    if (!f) return NULL;
    f->withz = 0;
    f->nzfptr = fopen(path, mode);
    if (!f->nzfptr) { free(f); return NULL; }
    return f;
}

int Xznzclose(znzFile *file)
{
    int rc = 0;
    if (file && *file) {
        if ((*file)->nzfptr) rc = fclose((*file)->nzfptr);
        free(*file);
        *file = NULL;
    }
    return rc;
}

size_t znzread(void *buf, size_t size, size_t nmemb, znzFile file)
{
    if (!file || !file->nzfptr) return 0;
    return fread(buf, size, nmemb, file->nzfptr);
}

size_t znzwrite(const void *buf, size_t size, size_t nmemb, znzFile file)
{
    if (!file || !file->nzfptr) return 0;
    return fwrite(buf, size, nmemb, file->nzfptr);
}

long znzseek(znzFile file, long offset, int whence)
{
    if (!file || !file->nzfptr) return -1;
    return fseek(file->nzfptr, offset, whence);
}

int znzrewind(znzFile stream)
{
    if (!stream || !stream->nzfptr) return -1;
    rewind(stream->nzfptr);
    return 0;
}

long znztell(znzFile file)
{
    if (!file || !file->nzfptr) return -1;
    return ftell(file->nzfptr);
}

int znzputs(const char *str, znzFile file)
{
    if (!file || !file->nzfptr) return -1;
    return fputs(str, file->nzfptr);
}
