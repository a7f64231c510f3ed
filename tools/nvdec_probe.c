/* tools/nvdec_probe.c -- does this box expose NVDEC through libnvcuvid?  (probe only)
 *   gcc tools/nvdec_probe.c -o tools/nvdec_probe -ldl && tools/nvdec_probe */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef struct {
    int eCodecType, eChromaFormat;
    unsigned int nBitDepthMinus8, reserved1[3];
    unsigned char bIsSupported, nNumNVDECs;
    unsigned short nOutputFormatMask;
    unsigned int nMaxWidth, nMaxHeight, nMaxMBCount;
    unsigned short nMinWidth, nMinHeight;
    unsigned char bIsHistogramSupported, nCounterBitDepth;
    unsigned short nMaxHistogramBins;
    unsigned int reserved3[10];
} CAPS;

int main(void)
{
    void *cu = dlopen("libcuda.so.1", RTLD_NOW), *nv = dlopen("libnvcuvid.so.1", RTLD_NOW);
    printf("libcuda %p libnvcuvid %p (%s)\n", cu, nv, nv ? "ok" : dlerror());
    if (!cu || !nv) return 1;
    int (*cuInit)(unsigned) = dlsym(cu, "cuInit");
    int (*cuDeviceGet)(int *, int) = dlsym(cu, "cuDeviceGet");
    int (*cuCtxCreate)(void **, unsigned, int) = dlsym(cu, "cuCtxCreate_v2");
    int (*cuDevicePrimaryCtxRetain)(void **, int) = dlsym(cu, "cuDevicePrimaryCtxRetain");
    int (*cuCtxPushCurrent)(void *) = dlsym(cu, "cuCtxPushCurrent_v2");
    int (*caps)(CAPS *) = dlsym(nv, "cuvidGetDecoderCaps");
    printf("cuvidGetDecoderCaps %p cuvidCreateVideoParser %p cuvidCreateDecoder %p cuvidMapVideoFrame64 %p\n", (void *)caps,
           dlsym(nv, "cuvidCreateVideoParser"), dlsym(nv, "cuvidCreateDecoder"), dlsym(nv, "cuvidMapVideoFrame64"));
    int dev; void *ctx;
    printf("cuInit %d\n", cuInit(0));
    printf("cuDeviceGet %d\n", cuDeviceGet(&dev, 0));
    printf("primary ctx %d\n", cuDevicePrimaryCtxRetain(&ctx, dev));
    printf("push %d\n", cuCtxPushCurrent(ctx));
    (void)cuCtxCreate;
    const int codecs[] = {4, 8, 10, 11, 5};
    const char *names[] = {"H264", "HEVC", "VP9", "AV1", "JPEG"};
    for (int i = 0; i < 5 && caps; i++) {
        CAPS c; memset(&c, 0, sizeof c);
        c.eCodecType = codecs[i]; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0;
        int r = caps(&c);
        printf("%s: ret %d supported %d nvdecs %d outmask 0x%x max %ux%u mb %u min %ux%u\n", names[i], r, c.bIsSupported, c.nNumNVDECs,
               c.nOutputFormatMask, c.nMaxWidth, c.nMaxHeight, c.nMaxMBCount, c.nMinWidth, c.nMinHeight);
    }
    return 0;
}
