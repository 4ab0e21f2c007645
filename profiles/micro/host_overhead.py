"""Profile the HOST side of batch.analyse_staged for 125 pairs: an engine that returns canned device results instantly."""
import sys, time, cProfile, pstats
import os; ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path[:0]=[ROOT, os.path.join(ROOT,'nightcore-to-flac-analyzer_b200'), os.path.join(ROOT,'tests')]
import numpy as np, torch
from fake_engine import FakeEngine
from nightcore_analyzer import _engine, batch
SR=22050
class NullEngine(FakeEngine):
    """shapes like the real thing for 180 s / 144 s pairs; values canned"""
    def window_energy_dev(self, audio, seg_off, seg_len):
        return torch.full((seg_off.numel(),), 0.01, dtype=torch.float64)
    def trim_bounds_dev(self, audio, seg_off, seg_len, top_db):
        return torch.from_numpy(np.stack([np.zeros(len(seg_len),np.int64), np.asarray(seg_len,np.int64)],1))
    def tempo_segments_dev(self, audio, seg_off, seg_len, start_bpm, hop, sr):
        n=len(seg_len)
        if hop==512:
            return None,None,None,torch.full((n,),21,dtype=torch.int32),torch.zeros((n,40),dtype=torch.int32),torch.full((n,),18,dtype=torch.int32)
        nb=330
        beats=torch.from_numpy((np.arange(nb,dtype=np.int32)*172)[None,:].repeat(n,0).copy())
        return None,None,None,torch.full((n,),172,dtype=torch.int32),beats,torch.full((n,),nb,dtype=torch.int32)
    def chroma_mean_dev(self, audio, seg_off, seg_len, sr, tuning_idx=None):
        return torch.rand((len(seg_len),12),dtype=torch.float64),None
    def cyclic_xcorr_dev(self, src, nc):
        return torch.full((src.shape[0],),4,dtype=torch.int32)
    def bootstrap(self, jobs, seed, n_boot, q_lo, q_hi, **kw):
        return np.tile(np.array([[1.25,1.24,1.26]]),(len(jobs),1)),None,None
eng=NullEngine()
_engine.get_engine=lambda device=None: eng
P=125
n_nc,n_src=3175200,3969000
lens=np.array([n_nc,n_src]*P,dtype=np.int64)
off=np.concatenate([[0],np.cumsum((lens+3)//4*4)[:-1]]).astype(np.int64)
audio=torch.zeros(4,dtype=torch.float32)   # never touched by the null engine
st=batch.StagedBatch(audio=audio,off=off,length=lens,sr=SR,h2d_bytes=0)
batch.analyse_staged(st)
t0=time.perf_counter()
for _ in range(3): batch.analyse_staged(st)
print('host time per 125-pair sub-batch: %.1f ms'%((time.perf_counter()-t0)/3*1e3))
cProfile.run('batch.analyse_staged(st)','/tmp/hp')
pstats.Stats('/tmp/hp').sort_stats('cumtime').print_stats(28)
