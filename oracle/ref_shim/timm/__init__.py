"""Test-infrastructure shim: the minimum of the `timm` API that the reference's
models/vlmo/{vlmo,vlmo_module}.py import (timm itself is not installed here).

Restated from timm's published behaviour (timm ~0.4.12-0.5.4, the era pinned by the
reference's misc/setup.sh); see SURVEY.md section 8(c). Only oracle/gen_golden.py puts
this directory on sys.path. Nothing in the product path imports it.
"""
