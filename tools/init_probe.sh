python - <<'PY'
import sys, os, subprocess, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from sigfish_b200 import synth, build as B
import helpers as H, gzip, shutil
d='/tmp/probe'; os.makedirs(d, exist_ok=True)
ids, sigs, sc = H.load_reads_npz(os.path.join(H.GOLDEN, "sp1_dna.npz"))
synth.write_blow5(d+'/r.blow5', ids, sigs, scalings=sc)
mean, stdv = synth.make_model(6); synth.write_model_file(d+'/m.txt', 6, mean, stdv)
with gzip.open(os.path.join(H.GOLDEN,'nCoV-2019.fa.gz'),'rb') as fi, open(d+'/ref.fa','wb') as fo: shutil.copyfileobj(fi,fo)
for i in range(2):
    t0=time.perf_counter()
    r=subprocess.run([B.CLI,'dtw',d+'/ref.fa',d+'/r.blow5','--kmer-model',d+'/m.txt','--verbose','5','--gpus','1'],capture_output=True,text=True)
    print('wall', time.perf_counter()-t0, r.returncode)
print(r.stderr[-1800:])
t0=time.perf_counter()
r=subprocess.run([H.REF_BIN,'dtw',d+'/ref.fa',d+'/r.blow5','--kmer-model',d+'/m.txt','-t','16'],capture_output=True,text=True)
print('ref wall', time.perf_counter()-t0)
PY
