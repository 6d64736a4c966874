"""Shared-memory arithmetic of the batch factorisation kernel (csrc/batch.cu, eeq_factor_ll_kernel), restated in Python:
the backward substitution stages two 8 KB chunks of L in whatever a molecule leaves free of the launch's shared memory --
behind b0 | b1 inside the block-column buffer and / or behind the molecule's own arrays.  The claim checked here is the
one the kernel's comment makes: with a launch sized for the largest class, two stages fit for EVERY augmented order, they
are 16-byte aligned, and they overlap neither each other nor anything the molecule still uses."""
KB, TS, CBW = 16, 17, 20
NPA_MAX = 224
STG = 2 * 8 * 64


def smem_doubles_ll(npa):  # batch.cu: smem_doubles_ll
    return (npa - KB if npa > KB else 2) * KB + KB * TS + KB * KB + 2 * CBW + 2 * KB + KB + 2 * KB + 8 + (2 * npa if npa <= 2 * KB else 0)


def stages(npa, dyn_doubles):
    own = smem_doubles_ll(npa)
    n_cb = ((npa - KB) * KB - 2 * npa) // STG if npa > 2 * KB else 0
    n_af = (dyn_doubles - own) // STG
    if n_cb + n_af < 2:
        return None
    cb0 = 0  # offset of the block-column buffer
    s0 = cb0 + 2 * npa if n_cb >= 1 else own
    s1 = cb0 + 2 * npa + STG if n_cb >= 2 else own + (0 if n_cb >= 1 else STG)
    return own, s0, s1


def test_two_stages_fit_for_every_order_in_a_full_size_launch():
    dyn = smem_doubles_ll(NPA_MAX)
    for npa in range(KB, NPA_MAX + 1, KB):
        st = stages(npa, dyn)
        assert st is not None, npa
        own, s0, s1 = st
        cb_end = (npa - KB) * KB if npa > KB else 2 * KB  # end of the block-column buffer
        busy = (0, 2 * npa) if npa > 2 * KB else (own - 2 * npa, own)  # b0 | b1
        for s in (s0, s1):
            assert s % 2 == 0  # 16-byte aligned
            assert s + STG <= dyn, (npa, s)
            inside_cb = s + STG <= cb_end and s >= 2 * npa and npa > 2 * KB
            behind = s >= own
            assert inside_cb or behind, (npa, s)  # never on the tiles T / Li, the scalars or b0 | b1
            assert s + STG <= busy[0] or s >= busy[1], (npa, s)
        assert abs(s0 - s1) >= STG, npa


def test_small_launches_fall_back_instead_of_overlapping():
    # launches sized for a smaller class (MCHRG_B200_FACTOR_GROUP < 13): either two stages fit or the kernel keeps its
    # direct loads; a stage never reaches beyond the launch
    for cap in range(2 * KB, NPA_MAX + 1, KB):
        dyn = smem_doubles_ll(cap)
        for npa in range(KB, cap + 1, KB):
            st = stages(npa, dyn)
            if st is None:
                continue
            _, s0, s1 = st
            assert max(s0, s1) + STG <= dyn, (cap, npa)
