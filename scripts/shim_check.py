import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np
import bauklank_audio_stretch_b200 as bs
from oracle.refdrive import *
lib = bs.load_library(os.environ.get('BSLIB') or None)
x = survey_clip(); sr=48000
xs = np.ascontiguousarray(x[:, :30000])
def cmpo(name, a, b):
    nz = np.nonzero(a.view(np.uint32)!=b.view(np.uint32))
    d = np.abs(a.astype(np.float64)-b.astype(np.float64))
    print(name, 'bit-identical' if nz[0].size==0 else 'DIFF n=%d first=%s max=%g'%(nz[0].size, nz[1].min(), d.max()), flush=True)
T=dict(tonality_hz=8000.0)
def both(name, fn):
    a = fn(PortEngine(5)); b = fn(bs.StretchEngine(seed=5, lib=lib))
    if isinstance(a, tuple): a, b = a[0], b[0]
    cmpo(name, a, b)
both('shim KA1', lambda e: stream_drive(e,x,sr,512,512,params=dict(semitones=0,**T)))
both('shim KA4', lambda e: kiosk_drive(e,x,sr,128000,0.75,preset='cheaper',params=dict(semitones=5,**T)))
both('shim KA6', lambda e: kiosk_drive(e,xs,sr,60000,0.5,params=dict(semitones=3,formant_semitones=4,formant_comp=True,formant_base_hz=200.0,**T)))
both('shim rng', lambda e: kiosk_drive(e,xs,sr,40000,0.3,params=dict(semitones=2,**T)))
both('shim stream split 480->512', lambda e: stream_drive(e,xs,sr,480,512,preset='cheaper',params=dict(semitones=1,**T)))
both('shim stream 100->900', lambda e: stream_drive(e,xs,sr,100,900,params=dict(semitones=0,**T)))
# parameter changes per quantum (Q4) in split mode
def pf(k, t): return dict(semitones=float((k//7)%13-6), rate=0.5+ (k%50)/40.0, formant_semitones=float((k//11)%5-2), formant_comp=bool(k%2))
both('shim param sweep split', lambda e: kiosk_drive(e,xs,sr,40000,1.0,preset='cheaper',params=dict(**T),param_fn=pf))
both('shim param sweep default', lambda e: kiosk_drive(e,xs,sr,40000,1.0,params=dict(**T),param_fn=pf))
# silence gate
z = np.zeros((2,40000),np.float32); z[:, :6000] = xs[:, :6000]; z[:, 30000:] = xs[:, :10000]
both('shim silence gate', lambda e: stream_drive(e,z,sr,512,512,params=dict(semitones=2,**T)))
