import sys
sys.path.insert(0, '/root/repo')
import _pkg
pkg = _pkg.load_package()
ctx = pkg.Context(0)
print(ctx.int32_peak())
