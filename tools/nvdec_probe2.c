/* second NVDEC probe: own context (cuCtxCreate) + try creating a decoder outright */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
typedef struct { unsigned long ulWidth, ulHeight, ulNumDecodeSurfaces; int CodecType, ChromaFormat; unsigned long ulCreationFlags, bitDepthMinus8, Reserved1[4];
  struct { short l,t,r,b; } display_area; int OutputFormat, DeinterlaceMode; unsigned long ulTargetWidth, ulTargetHeight, ulNumOutputSurfaces; void *vidLock;
  struct { short l,t,r,b; } target_rect; unsigned long Reserved2[5]; } CI;
int main(void)
{
    void *cu = dlopen("libcuda.so.1", RTLD_NOW), *nv = dlopen("libnvcuvid.so.1", RTLD_NOW);
    if (!cu || !nv) { printf("dlopen failed\n"); return 1; }
    int (*cuInit)(unsigned) = dlsym(cu, "cuInit");
    int (*cuDeviceGet)(int *, int) = dlsym(cu, "cuDeviceGet");
    int (*cuCtxCreate)(void **, unsigned, int) = dlsym(cu, "cuCtxCreate_v2");
    int (*cuGetErrorName)(int, const char **) = dlsym(cu, "cuGetErrorName");
    int (*create)(void **, CI *) = dlsym(nv, "cuvidCreateDecoder");
    int (*destroy)(void *) = dlsym(nv, "cuvidDestroyDecoder");
    int (*lockc)(void **, void *) = dlsym(nv, "cuvidCtxLockCreate");
    int dev; void *ctx = 0;
    printf("cuInit %d\n", cuInit(0)); printf("cuDeviceGet %d\n", cuDeviceGet(&dev, 0));
    printf("cuCtxCreate %d\n", cuCtxCreate(&ctx, 0, dev));
    void *lk = 0; if (lockc) printf("cuvidCtxLockCreate %d\n", lockc(&lk, ctx));
    for (int codec = 4; codec <= 8; codec += 4) {
        CI ci; memset(&ci, 0, sizeof ci);
        ci.ulWidth = 1920; ci.ulHeight = 1088; ci.ulNumDecodeSurfaces = 8; ci.CodecType = codec; ci.ChromaFormat = 1;
        ci.ulCreationFlags = 4; ci.OutputFormat = 0; ci.DeinterlaceMode = 0; ci.ulTargetWidth = 1920; ci.ulTargetHeight = 1080; ci.ulNumOutputSurfaces = 2;
        ci.display_area.r = 1920; ci.display_area.b = 1080; ci.vidLock = lk;
        void *dec = 0; int r = create(&dec, &ci); const char *n = "?"; if (cuGetErrorName) cuGetErrorName(r, &n);
        printf("cuvidCreateDecoder codec %d -> %d (%s) dec=%p\n", codec, r, n, dec);
        if (dec) destroy(dec);
    }
    return 0;
}
