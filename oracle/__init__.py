"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the Viterbi hot path of drwangxian/viterbi_spl.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this package; the product (``viterbi_spl_b200``) never does and fails loudly without its CUDA library.

Modules
-------
np_oracle   line-for-line NumPy restatement of ``imm/tf_viterbi.py:75-109`` (+ the family wrappers)
c_oracle    ctypes binding of ``viterbi_oracle.c`` (same algorithm, C, pthread batch) for large cases
ref_loader  AST loader that executes the reference's OWN functions from ``/root/reference`` (this
            container only; used to pin the two restatements and to generate ``tests/golden``)
fb_oracle   float64 scaled forward-backward (no reference implementation exists: parity unpinned)
"""
