"""the builder's lambda-graph configuration (host logic, CPU only)."""


def test_builder_defaults_and_result_k(sfb):
    b = sfb.LambdaGraphBuilder()
    # surfface-pipeline/src/builder.rs:105-111
    assert (b.lambda_eps, b.lambda_k, b.lambda_topk, b.lambda_p, b.lambda_sigma, b.normalise, b.sparsity_check) == (1e-3, 6, 3, 2.0, None, False, False)
    # define_result_k (:785-793): k <= 5 -> topk 3; 5 < k < 10 -> topk 4; larger k leaves the user's topk
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 3, 9, 2.0).graph_params().topk == 3
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 5, 9, 2.0).graph_params().topk == 3
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 6, 9, 2.0).graph_params().topk == 4
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 9, 9, 2.0).graph_params().topk == 4
    assert sfb.LambdaGraphBuilder().with_lambda_graph(0.5, 10, 9, 2.0).graph_params().topk == 9
    gp = sfb.LambdaGraphBuilder().with_lambda_graph(0.25, 16, 16, 1.5, sigma_override=0.3).graph_params()
    assert gp == sfb.GraphParams(0.25, 16, 16, 1.5, 0.3, False, False)
    assert sfb.LambdaGraphBuilder().graph_params() == sfb.GraphParams(1e-3, 6, 4, 2.0, None, False, False)   # the default k = 6 lands on topk 4


def test_jl_dimension_host_function_matches_the_reference_table(sfb, oracle):
    """sfb_compute_jl_dimension is host arithmetic (no device): every exact value of src_legacy/tests/
    test_reduction.rs:193-562 through the C ABI, and agreement with the oracle's restatement on a grid."""
    jl = sfb.compute_jl_dimension
    assert [jl(100, 16, 0.3), jl(1000, 8, 0.1), jl(50, 31, 0.2), jl(10, 1, 0.5)] == [16, 8, 31, 1]
    assert jl(10, 100, 0.3) == 100 and jl(10, 50, 0.3) == 50 and jl(2, 1000, 0.9) == 32 and jl(2, 20, 0.9) == 20
    assert jl(1000, 512, 0.1) == 512 and jl(1, 100, 0.1) == 32 and jl(1, 10, 0.1) == 10
    for n in (0, 1, 2, 17, 100, 5000, 10**6):
        for d in (1, 31, 32, 64, 384, 2048, 2049, 5000, 100_000):
            for e in (0.05, 0.15, 0.3, 0.5, 0.9):
                assert jl(n, d, e) == oracle.jl_dimension(n, d, e), (n, d, e)
                assert jl(n, d, e, core=True) == oracle.jl_dimension(n, d, e, core=True), (n, d, e)
