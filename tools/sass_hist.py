"""SASS mnemonic histogram of one compiled translation unit (evidence that the tcgen05 / TMA path is what ships):
    python tools/sass_hist.py poet_b200/build/gemm_tc.o > profiles/r02_sass_gemm_tc.txt
Lists, per kernel, the instruction count and the Blackwell-specific mnemonics (UTC*MMA = tcgen05.mma, LDTM/STTM =
tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce, UBLKCP/UBLKPF = bulk copy / prefetch, RED/REDG =
global reductions, HMMA = legacy mma.sync, SYNCS = mbarrier) and then the full histogram of the object."""
import collections
import re
import subprocess
import sys

KEY = re.compile(r"^(UTC\w*MMA|UTCBAR|UTCATOM\w*|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTMAPF|UBLKCP|UBLKPF|UBLKRED|RED|REDG|ATOM\w*|HMMA|SYNCS|USETMAXREG|"
                 r"LDGSTS|F2FP|MUFU|REDUX|LDSM|STSM|BAR|ACQBULK|UCGABAR\w*|ELECT|LDS|STS|LDG|STG|SHFL)$")


def main():
    obj = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    total = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur][op] += 1
            total[op + mods] += 1
    print(f"# cuobjdump -sass {obj}: {len(kernels)} kernels, {sum(sum(c.values()) for c in kernels.values())} instructions")
    for name, c in kernels.items():
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        short = (dem.split(">(")[0] + ">" if ">(" in dem else re.sub(r"\(.*", "", dem))[:170]
        keys = {k: v for k, v in c.items() if KEY.match(k)}
        print(f"\n## {short}\n   instructions={sum(c.values())}  " + "  ".join(f"{k}={v}" for k, v in sorted(keys.items())))
    print("\n# object-wide histogram (mnemonic with modifiers, count >= 4)")
    for k, v in total.most_common():
        if v >= 4 and KEY.match(k.split(".")[0]):
            print(f"{k:48s} {v}")


if __name__ == "__main__":
    main()
