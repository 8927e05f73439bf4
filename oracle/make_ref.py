"""Mirror the reference's hot-path sources into oracle/_ref/ so that the UNMODIFIED reference can run on the GPU box.

    python oracle/make_ref.py          # run in the build container (needs /root/reference)

The GPU box has no /root/reference; gpurun ships /root/repo only. oracle/_ref/ is git-ignored (the reference's
sources never enter this repository's history) but not gpurun-ignored, so the mirrored files travel with the
snapshot like a built .so. What is mirrored: `models/` (build.py, vlmo/{vlmo,vlmo_module,objectives,heads}.py,
modeling_discrete_vae.py) and `dall_e/` (imported by objectives.py at module load) — the import closure of
`models.build.build_model`, byte for byte. `oracle/ref_run.py` imports them through the timm shim
(oracle/ref_shim). Test infrastructure only: the product never imports any of it.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('MOME_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'oracle', '_ref')
TREES = ('models', 'dall_e')


def main():
    if not os.path.isdir(os.path.join(SRC, 'models', 'vlmo')):
        print(f'make_ref: {SRC} not present (GPU box?) - keeping whatever is in {DST}')
        return 0
    manifest = []
    for tree in TREES:
        for dirpath, _, files in os.walk(os.path.join(SRC, tree)):
            for f in sorted(files):
                if not f.endswith('.py'):
                    continue
                src = os.path.join(dirpath, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(src, 'rb') as fh:
                    manifest.append(f'{hashlib.sha256(fh.read()).hexdigest()}  {rel}')
    with open(os.path.join(DST, 'MANIFEST.sha256'), 'w') as fh:
        fh.write('\n'.join(manifest) + '\n')
    print(f'make_ref: mirrored {len(manifest)} files into {DST}')
    return 0


if __name__ == '__main__':
    sys.exit(main())
