import importlib.util, sys, time, os
ROOT='/root/repo'
spec = importlib.util.spec_from_file_location("cuda_lbfgs_b200", os.path.join(ROOT, "cuda-lbfgs_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); sys.modules["cuda_lbfgs_b200"]=pkg; spec.loader.exec_module(pkg)
L=pkg.lib()
n=int(sys.argv[1]) if len(sys.argv)>1 else 100_000_000
x0=pkg.PinnedArray(n); out=pkg.PinnedArray(n)
pkg.x0_uniform(n,-2,2,out=x0.array)
for graph in (0,1,1):
    p=pkg.default_params("par", line_search="wolfe", m=10, max_iterations=10**9, tolerance=0.0, use_graph=graph, direction="compact")
    L.lbfgsb200_device_sync(); t=[time.perf_counter()]
    s=pkg.Solver("rosenbrock", n, p); L.lbfgsb200_device_sync(); t.append(time.perf_counter())
    s.set_x0(x0.array); L.lbfgsb200_device_sync(); t.append(time.perf_counter())
    s.iterate(1); L.lbfgsb200_device_sync(); t.append(time.perf_counter())
    s.iterate(41); L.lbfgsb200_device_sync(); t.append(time.perf_counter())
    s.x(out=out.array); t.append(time.perf_counter())
    s.destroy(); L.lbfgsb200_device_sync(); t.append(time.perf_counter())
    names=["create","set_x0","first iterate(1)","iterate(41)","get_x","destroy"]
    print("graph=%d"%graph, {k: round((t[i+1]-t[i])*1e3,2) for i,k in enumerate(names)}, "total ms", round((t[-1]-t[0])*1e3,1))
