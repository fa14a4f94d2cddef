import json, sys, glob, os
libs = sys.argv[1:] or sorted({os.path.basename(f)[3:].rsplit('_', 1)[0] for f in glob.glob('gpurun_out/ab_*_unshared.json')})
for lib in libs:
    for v in ("unshared", "shared", "b20"):
        try:
            d = json.loads(open(f"gpurun_out/ab_{lib}_{v}.json").read().strip().splitlines()[-1])
            print("%-10s %-9s value %7.2f G  enc %7.2f ms  agg %5.2f  dec %6.3f  prf %5.1f G" % (
                lib, v, d["value"] / 1e9, d["phases"]["encode_encrypt_ms"], d["phases"]["aggregate_ms"],
                d["phases"]["decrypt_decode_ms"], d["roofline_prf"]["achieved"]))
        except Exception as e:
            print(lib, v, "ERR", e)
