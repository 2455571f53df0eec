import json, sys
d = json.load(open(sys.argv[1]))
print('kernel ms/step %.3f   step ms %.3f' % (d['kernel_ms_per_step'], d['step_ms']))
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms']):
    print('%-44s n/step %4d  ms/step %7.3f  avg us %7.1f  share %5.1f%%  %8.1f TF/s %8.1f GB/s' % (
        k, v['launches'] // 3, v['ms_per_step'], 1e3 * v['ms'] / v['launches'], 100 * v['share'], v['tflops'], v['gbs']))
