"""CPU oracle for the EBEN bandwidth-extension training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vibravox_b200/`` may import this
package: it is the checker that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` run beside the
CUDA path, never the thing that is shipped or measured as the product.

Pinning status (see DESIGN.md "Oracle"):
  * generator / discriminators / PQMF / feature-matching / hinge / training
    step: pinned against the reference's own modules imported from
    /root/reference in the build container (``oracle/make_goldens.py``), with the
    outputs committed under ``tests/golden/``.
  * multi-resolution STFT loss: the arithmetic lives in the third-party package
    ``auraloss`` (unpinned in the reference's pyproject.toml:21, not vendored,
    not installable offline).  It is restated from the published algorithm
    (auraloss 0.4.0 ``freq.py`` / ``perceptual.py``) -> **parity unpinned** at the
    dependency level for that one loss.
"""
