#!/usr/bin/env python3
"""Golden vectors for off-target site extraction, produced by the UNMODIFIED reference tool
(/root/reference/src/crackling/utils/extractOfftargets.py) on small, deliberately messy FASTA inputs:
mixed case, N runs and IUPAC codes, LF / CRLF / lone-CR line ends, surrounding blanks, '>' inside a line,
very short lines, records shorter than a window, repeated headers, a headerless file.

Commits the inputs (tests/golden/extract/<case>/input*.fa) and, per case, the sha256 / line count / first and
last lines of what the tool wrote.  Run in the build container (needs /root/reference)."""
import hashlib, json, os, pathlib, subprocess, sys, tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import extract  # noqa: E402


def rand_seq(rng, n, alphabet="ACGT", p=None):
    return "".join(rng.choice(list(alphabet), size=n, p=p))


def wrap(seq, width, eol="\n"):
    return "".join(seq[i:i + width] + eol for i in range(0, len(seq), width))


def pam_rich(rng, n):
    """Random sequence with extra GG / CC so that both patterns fire often, plus planted exact repeats."""
    s = list(rand_seq(rng, n))
    for i in rng.integers(0, n - 2, n // 12):
        s[i:i + 2] = "GG" if rng.random() < 0.5 else "CC"
    unit = s[100:160]
    for i in rng.integers(200, n - 100, 6):
        s[i:i + 60] = unit
    return "".join(s)


def case_single(rng):
    recs = []
    a = pam_rich(rng, 14000)
    recs.append(">chr1 plain LF, width 60\n" + wrap(a, 60))
    b = pam_rich(rng, 9000)
    b = b[:3000] + "N" * 137 + b[3000:6000].lower() + "RYKMSWBDHVN" + b[6000:]
    recs.append(">chr2 soft-masked, N run, IUPAC; CRLF\r\n" + wrap(b, 70, "\r\n"))
    c = pam_rich(rng, 5000)
    recs.append(">chr3 lone CR line ends\r" + wrap(c, 50, "\r"))
    d = pam_rich(rng, 4000)
    lines = [d[i:i + 40] for i in range(0, len(d), 40)]
    for k in range(0, len(lines), 7):
        lines[k] = "  \t" + lines[k] + " \t "
    lines[5] = lines[5][:20] + " " + lines[5][20:]          # a blank inside a line stays and breaks windows
    lines[9] = lines[9][:10] + ">" + lines[9][10:]          # '>' inside a line is just a character
    lines[11] = lines[11][:7] + "12*-" + lines[11][7:]
    recs.append("   >chr4 indented header, padded lines\n" + "".join(l + "\n" for l in lines))
    e = pam_rich(rng, 700)
    cuts = np.r_[0, np.cumsum(rng.integers(1, 6, 400))]
    recs.append(">chr5 one to five characters per line\n" + "".join(e[i:j] + "\n" for i, j in zip(cuts[:-1], cuts[1:]) if i < len(e)))
    recs.append(">tiny shorter than a window\nACGTACGTACGTACGTACGTGG\n")
    recs.append(">edge window ends exactly at the record end\n" + "ACGTTGCAACGTTGCAACGTAGG\n")
    f = pam_rich(rng, 6000)
    recs.append(">chr8 no newline at the end of the file\n" + wrap(f, 80)[:-1])
    return {"input0.fa": "".join(recs).encode()}


def case_several(rng):
    a = pam_rich(rng, 8000)
    files = {"input0.fa": wrap(a, 61).encode()}                                     # no header at all
    b1, b2, b3 = pam_rich(rng, 3000), pam_rich(rng, 3500), pam_rich(rng, 2500)
    files["input1.fa"] = (">dup\n" + wrap(b1, 60) + ">other\n" + wrap(b2.lower(), 60) + "\n\n>dup\n" + wrap(b3, 60)).encode()
    c = pam_rich(rng, 5000)
    lines = [c[i:i + 45] for i in range(0, len(c), 45)]
    lines[3] = "  " + lines[3]                                                        # leading blanks stay on this path
    lines[8] = lines[8] + "  \t"
    files["input2.fa"] = (">crlf\r\n" + "".join(l + "\r\n" for l in lines) + " >not a header\r\n" + wrap(pam_rich(rng, 900), 30, "\r\n")).encode()
    return files


def run_tool(files: dict) -> bytes:
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for name, data in files.items():
            pathlib.Path(tmp, name).write_bytes(data)
            paths.append(str(pathlib.Path(tmp, name)))
        env = dict(os.environ, PYTHONPATH="/root/reference/src")
        subprocess.run([sys.executable, "-m", "crackling.utils.extractOfftargets", f"{tmp}/sites.txt", *paths],
                       check=True, env=env, stdout=subprocess.DEVNULL)
        return pathlib.Path(tmp, "sites.txt").read_bytes()


def main():
    for name, maker, seed in (("single_messy", case_single, 101), ("several_files", case_several, 202)):
        files = maker(np.random.default_rng(seed))
        out = run_tool(files)
        mine = extract.extract_from_inputs([files[k] for k in sorted(files)])
        assert mine == out, f"{name}: oracle restatement differs from the reference tool ({len(mine)//21} vs {len(out)//21} lines)"
        d = HERE / "extract" / name
        d.mkdir(parents=True, exist_ok=True)
        for k, v in files.items():
            (d / k).write_bytes(v)
        lines = out.decode().splitlines()
        (d / "expected.json").write_text(json.dumps({
            "inputs": sorted(files), "sites_sha256": hashlib.sha256(out).hexdigest(), "sites_lines": len(lines),
            "distinct": len(set(lines)), "head": lines[:5], "tail": lines[-5:]}, indent=1))
        print(name, len(lines), "sites,", len(set(lines)), "distinct")


if __name__ == "__main__":
    main()
