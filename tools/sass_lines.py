#!/usr/bin/env python
"""Static code size per source line / routine of one kernel: nvdisasm -g <cubin> | python tools/sass_lines.py <mangled-substring>"""
import collections
import re
import sys

from ncu_lines import routines  # noqa: E402  (same directory)


def main():
    want = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    on, cur = False, None
    cnt = collections.Counter()
    paths = {}
    for l in sys.stdin:
        if l.startswith(".text."):
            on = want in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            f = m.group(1)
            paths[f.split("/")[-1]] = f
            cur = (f.split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            cnt[cur] += 1
    tot = sum(cnt.values())
    print("total SASS instructions", tot, f"({tot * 16 / 1024:.0f} KB)")
    fn = collections.Counter()
    for (f, ln), v in ((k, v) for k, v in cnt.items() if k):
        name = "?"
        for start, nm in routines(paths[f]):
            if start <= ln:
                name = nm
        fn[f + ":" + name] += v
    for k, v in fn.most_common(top):
        print(f"  {k:45s} {v:6d} {100 * v / tot:5.1f}%")
    print("top lines")
    for k, v in cnt.most_common(top):
        print("  ", k, v)


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    main()
