import sys, time
sys.path.insert(0, 'juliabugs.jl_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
exec(open('scripts/jit_probe_rats.py').read().split('a = run("interpreter ", 2)')[0])
a = run("interpreter ", 3)
