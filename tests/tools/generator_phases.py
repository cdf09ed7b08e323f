import subprocess, sys, time, os, random, tempfile
B='/root/repo/integration/_build'
rnd=random.Random(20482048); m=2048
r=2**(m-1)+1+rnd.randrange(2**(m-1)-1); d=r//2+rnd.randrange(r//2)
t=tempfile.mkdtemp(); os.makedirs(t+'/distributions')
cmd=[B+'/minimpirun','-np','2',B+'/gpu/generate_distribution','-exp',str(d),str(r),'-dim','256','2048','1']
t0=time.time()
p=subprocess.Popen(['stdbuf','-oL']+cmd,cwd=t,stdout=subprocess.PIPE,text=True)
marks={}
n=0
for line in p.stdout:
    n+=1
    for key in ('Processing slice: 1 /','Stopping node','Waiting for all export','Sorting the slices','Exporting distribution information','Exporting collapsed distribution to "distributions/collapsed-d','Exporting the distribution to','Finished exporting'):
        if key in line and key not in marks: marks[key]=time.time()-t0
p.wait()
print('total',time.time()-t0)
for k,v in marks.items(): print(f'{v:8.2f}s  {k}')
