"""Shared-memory-port and L2->SM accounting of the 3xTF32 tcgen05 GEMM tiles ("The 3xTF32 GEMMs ...").

Per k-step of 8 (one tcgen05.mma kind::tf32 instruction each for hi.hi, lo.hi, hi.lo) an SM
  * reads the A and B operand slices of the instruction from shared memory, three times (one per product);
  * receives the hi and lo copies of both slices from TMA once (shared-memory writes, and L2 -> SM bytes);
against a 128 B/clk shared-memory port and a tensor pipe that retires 2048 tf32 MACs per clock and SM
(2.25 PFLOP/s dense bf16 nominal / 2 for tf32 / 148 SMs / ~1.9 GHz ~= 4096 FLOP/clk/SM).

    python scripts/gemm_port_model.py
"""
PORT = 128.0          # B/clk/SM shared memory
MACS = 2048.0         # tf32 MAC/clk/SM
SMS, GHZ = 148, 1.9


def tile(name, m_rows, n_rows_local, n_cols_mma, in_kernel_lo=False):
    """m_rows: A rows staged per SM; n_rows_local: B rows staged per SM (half the tile on a CTA pair);
    n_cols_mma: N of the instruction (what the SM's tensor core multiplies its m_rows by)."""
    k = 8
    a, b = m_rows * k * 4, n_rows_local * k * 4                      # bytes of one operand slice per k-step
    clk = 3 * m_rows * n_cols_mma * k / MACS                          # tensor cycles of the three products
    reads = 3 * (a + b)                                               # operand reads of the three instructions
    fill = 2 * (a + b) if not in_kernel_lo else (a + b) + 2 * (a + b)  # TMA hi+lo | TMA fp32 + converter read + lo write
    l2 = 2 * (a + b) if not in_kernel_lo else (a + b)
    port = (reads + fill) / clk
    print("%-34s reads %5.1f + fill %5.1f = %5.1f B/clk  (port %3.0f %%: tensor pipe <= %3.0f %%)   L2->SM at 100 %% tensor: %4.1f TB/s" % (
        name, reads / clk, fill / clk, port, 100 * port / PORT, min(100.0, 100 * PORT / port), l2 / clk * SMS * GHZ * 1e9 / 1e12))


tile("single-SM 128 x 160 (gP, k-major)", 128, 160, 160)
tile("single-SM 128 x 256 (gT, mn-major)", 128, 256, 256)
tile("CTA pair 256 x 256 (forward)", 128, 128, 256)
tile("CTA pair, lo derived in-kernel", 128, 128, 256, in_kernel_lo=True)
tile("single-SM 128 x 160, lo in-kernel", 128, 160, 160, in_kernel_lo=True)
