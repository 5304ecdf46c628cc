import os, sys, time
sys.path.insert(0,'/root/repo')
import unicycler_b200 as ub
os.environ['UNICYCLER_B200_SEED']='7'
for rep in range(2):
    t0=time.time(); r=ub.get_random_sequence_alignment_mean_and_std_dev(100,25000,(3,-6,-5,-2)); print(r, '%.1f ms'%((time.time()-t0)*1e3), ub.last_stats(), flush=True)
