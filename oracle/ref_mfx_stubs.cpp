/*
 * TEST INFRASTRUCTURE ONLY.  Link-time stand-ins for the Intel Media SDK dispatcher entry
 * points that /root/reference/intel_dec/intel_dec.cpp and intel_enc/intel_enc.cpp reference
 * (the reference ships only Windows import libs, intel_sdk/lib/).  The checker never reaches
 * them: it calls the CPU surface-copy functions only.  Every stub reports MFX_ERR_UNSUPPORTED.
 */
extern "C" {
#define JMREF_MFX_STUB(name) __attribute__((visibility("default"))) int name(...) { return -3; }
JMREF_MFX_STUB(MFXInit)
JMREF_MFX_STUB(MFXClose)
JMREF_MFX_STUB(MFXQueryIMPL)
JMREF_MFX_STUB(MFXQueryVersion)
JMREF_MFX_STUB(MFXVideoCORE_SyncOperation)
JMREF_MFX_STUB(MFXVideoDECODE_Close)
JMREF_MFX_STUB(MFXVideoDECODE_DecodeFrameAsync)
JMREF_MFX_STUB(MFXVideoDECODE_DecodeHeader)
JMREF_MFX_STUB(MFXVideoDECODE_Init)
JMREF_MFX_STUB(MFXVideoDECODE_Query)
JMREF_MFX_STUB(MFXVideoDECODE_QueryIOSurf)
JMREF_MFX_STUB(MFXVideoENCODE_Close)
JMREF_MFX_STUB(MFXVideoENCODE_EncodeFrameAsync)
JMREF_MFX_STUB(MFXVideoENCODE_GetVideoParam)
JMREF_MFX_STUB(MFXVideoENCODE_Init)
JMREF_MFX_STUB(MFXVideoENCODE_Query)
JMREF_MFX_STUB(MFXVideoENCODE_QueryIOSurf)
JMREF_MFX_STUB(MFXVideoUSER_Load)
JMREF_MFX_STUB(MFXVideoUSER_UnLoad)
}
