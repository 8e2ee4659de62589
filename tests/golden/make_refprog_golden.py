"""Golden outputs for tests/test_gpu_refprogs.py: run the reference's own
example and test programs, built from the UNMODIFIED reference sources
(oracle/_ref, see oracle/Makefile), and record their stdout.  Run in the build
container (needs /root/reference to have been compiled):

    make -C oracle ref && python tests/golden/make_refprog_golden.py

The GPU test runs the SAME programs compiled against this repository's
headers and lib_fftwpp.so (tests/refprogs/Makefile) with the same arguments."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref")

EXAMPLES = ["exampleconv", "exampleconv2", "exampleconv3", "exampleconvh", "exampleconvh2",
            "exampleconvh3", "exampleconvr", "exampleconvr2", "exampleconvr3", "cexample"]


def cases():
    out = [(p, []) for p in EXAMPLES]
    # accuracy drivers: -E (error vs direct convolution) and -O (print result);
    # forced (m,D,I) as the reference's tests/tests.py passes them
    conv = {
        "hybridconv": [["-L8", "-M16"], ["-L8", "-M16", "-m4", "-D1", "-I0"],
                       ["-L8", "-M16", "-m8", "-D2", "-I0"], ["-L7", "-M20", "-m4", "-D1", "-I1"],
                       ["-L5", "-M10", "-m2", "-D1", "-I0"], ["-L16", "-M32", "-m16", "-D1", "-I1"],
                       ["-L8", "-M24", "-m8", "-D1", "-I0", "-c"]],
        "hybridconvh": [["-L8", "-M12"], ["-L8", "-M12", "-m4", "-D2", "-I0"],
                        ["-L7", "-M12", "-m12", "-D1", "-I0"], ["-L16", "-M24", "-m8", "-D2", "-I1"]],
        "hybridconvr": [["-L8", "-M16"], ["-L8", "-M16", "-m8", "-D1", "-I0"],
                        ["-L8", "-M16", "-m4", "-D1", "-I0"], ["-L5", "-M15", "-m5", "-D1", "-I1"]],
        "hybridconv2": [["-L8", "-M16"], ["-Lx=8", "-Ly=4", "-Mx=16", "-My=8", "-mx=8", "-my=4",
                                           "-Dx=1", "-Dy=1", "-Ix=1", "-Iy=0"],
                        ["-L4", "-M8", "-m2", "-D1", "-I1"], ["-L8", "-M16", "-m8", "-D1", "-I1", "-Sx=10"]],
        "hybridconvh2": [["-L8", "-M12"], ["-L8", "-M12", "-m4", "-Dx=1", "-Dy=2", "-I1"],
                         ["-Lx=7", "-Ly=8", "-Mx=11", "-My=12"]],
        "hybridconvr2": [["-L8", "-M16"], ["-L8", "-M16", "-m8", "-D1", "-I1"],
                         ["-Lx=6", "-Ly=8", "-Mx=12", "-My=16"]],
        "hybridconv3": [["-L4", "-M8"], ["-L4", "-M8", "-m4", "-D1", "-I1"],
                        ["-Lx=4", "-Ly=6", "-Lz=8", "-Mx=8", "-My=12", "-Mz=16"]],
        "hybridconvh3": [["-L4", "-M6"], ["-L8", "-M12", "-m4", "-Dx=1", "-Dy=1", "-Dz=2", "-I1"],
                         ["-Lx=5", "-Ly=4", "-Lz=8", "-Mx=8", "-My=6", "-Mz=12"]],
        "hybridconvr3": [["-L4", "-M8"], ["-L8", "-M16", "-m8", "-D1", "-I1"],
                         ["-L8", "-M16", "-mx=8", "-my=4", "-mz=4", "-D1", "-I1"],
                         ["-Lx=4", "-Ly=6", "-Lz=8", "-Mx=8", "-My=12", "-Mz=16"]],
    }
    for prog, lst in conv.items():
        for args in lst:
            out.append((prog, args + ["-E", "-O", "-T1"]))
    # forward/backward identity drivers (tests/hybrid.cc, hybridh.cc, hybridr.cc)
    pads = {
        "hybrid": [["-L8", "-M16", "-m4", "-D1", "-I0"], ["-L8", "-M16", "-m8", "-D2", "-I0"],
                   ["-L8", "-M32", "-m8", "-D4", "-I0"], ["-L7", "-M20", "-m4", "-D1", "-I1"],
                   ["-L8", "-M16", "-m8", "-D1", "-I1", "-C3", "-S4"],
                   ["-L8", "-M24", "-m4", "-D2", "-I0", "-c"], ["-L12", "-M36", "-m3", "-D1", "-I0"]],
        "hybridh": [["-L8", "-M12", "-m4", "-D2", "-I0"], ["-L8", "-M12", "-m12", "-D1", "-I0"],
                    ["-L16", "-M24", "-m4", "-D2", "-I1"], ["-L8", "-M12", "-m4", "-D2", "-I1", "-C2"]],
        "hybridr": [["-L8", "-M16", "-m8", "-D1", "-I0"], ["-L8", "-M16", "-m4", "-D1", "-I0"],
                    ["-L8", "-M32", "-m8", "-D1", "-I1"], ["-L8", "-M16", "-m8", "-D1", "-I1", "-C2", "-S2"],
                    ["-L16", "-M32", "-m4", "-D1", "-I0"]],
    }
    for prog, lst in pads.items():
        for args in lst:
            out.append((prog, args + ["-T1"]))
    return out


def main():
    rec = []
    for prog, args in cases():
        exe = os.path.join(REF, prog)
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        rec.append({"prog": prog, "args": args, "rc": r.returncode, "stdout": r.stdout})
        print(prog, " ".join(args), "rc", r.returncode, len(r.stdout), "bytes")
    with open(os.path.join(HERE, "refprogs.json"), "w") as fh:
        json.dump(rec, fh, indent=0)


if __name__ == "__main__":
    main()
