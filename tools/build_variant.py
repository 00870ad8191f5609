import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from adapt_b200.build import build
name, flags = sys.argv[1], sys.argv[2:]
print(build(force=True, extra_flags=flags, out=f'/root/repo/adapt_b200/lib/{name}.so'))
