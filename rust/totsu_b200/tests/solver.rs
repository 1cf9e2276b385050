// Backend conformance test: the same 1-variable SDP every backend of the reference runs
// (totsu_core/tests/solver.rs:13-54, totsu_f64lapack/tests/solver.rs:15-56, totsu_f32cuda/tests/solver.rs:14-55),
// instantiated for B200.  Expected: x[0] = -2 +- 1e-3.  The C++ mirror of this test is
// tests/test_solver_gpu.py::test_backend_conformance_sdp in the repository.
use float_eq::assert_float_eq;
use totsu_b200::{ProductCone, B200, TB_CONE_PSD};
use totsu_core::solver::{Operator, Solver};
use totsu_core::{ConePSD, MatOp, MatType};

type La = B200;
type AMatOp<'a> = MatOp<'a, La>;
type AConePSD<'a> = ConePSD<'a, La>;
type ASolver = Solver<La>;

fn problem<'a>() -> (AMatOp<'a>, AMatOp<'a>, AMatOp<'a>) {
    let op_c = AMatOp::new(MatType::General(1, 1), &[1.]);
    // vec of [[0,-1],[-1,-3]] and [[1,0],[0,10]]: upper triangle by columns, off-diagonals scaled by sqrt(2)
    let op_a = AMatOp::new(MatType::General(3, 1), &[0., -1. * 1.41421356, -3.]);
    let op_b = AMatOp::new(MatType::General(3, 1), &[1., 0. * 1.41421356, 10.]);
    (op_c, op_a, op_b)
}

#[test]
fn test_solver() {
    let _ = env_logger::builder().is_test(true).try_init();
    let (op_c, op_a, op_b) = problem();
    let s = ASolver::new().par(|p| p.max_iter = Some(100_000));
    let mut cone_w = vec![0.; AConePSD::query_worklen(op_a.size().0)];
    let cone = AConePSD::new(&mut cone_w, s.par.eps_zero);
    let mut work = vec![0.; ASolver::query_worklen(op_a.size())];
    let rslt = s.solve((op_c, op_a, op_b, cone, &mut work)).unwrap();
    assert_float_eq!(rslt.0[0], -2., abs_all <= 1e-3);
}

#[test]
fn test_solver_fused_cone() {
    let (op_c, op_a, op_b) = problem();
    let s = ASolver::new().par(|p| p.max_iter = Some(100_000));
    let cone = ProductCone::new(&[(TB_CONE_PSD, 3)], s.par.eps_zero);
    let mut work = vec![0.; ASolver::query_worklen(op_a.size())];
    let rslt = s.solve((op_c, op_a, op_b, cone, &mut work)).unwrap();
    assert_float_eq!(rslt.0[0], -2., abs_all <= 1e-3);
}

// The f64 twin of the backend: same conformance problem, the reference's default eps_acc = 1e-6 converges in it.
#[test]
fn test_solver_f64() {
    use totsu_b200::B200F64;
    type Op<'a> = MatOp<'a, B200F64>;
    let op_c = Op::new(MatType::General(1, 1), &[1.]);
    let op_a = Op::new(MatType::General(3, 1), &[0., -1. * 1.41421356, -3.]);
    let op_b = Op::new(MatType::General(3, 1), &[1., 0. * 1.41421356, 10.]);
    let s = Solver::<B200F64>::new().par(|p| p.max_iter = Some(100_000));
    let mut cone_w = vec![0.; ConePSD::<B200F64>::query_worklen(op_a.size().0)];
    let cone = ConePSD::<B200F64>::new(&mut cone_w, s.par.eps_zero);
    let mut work = vec![0.; Solver::<B200F64>::query_worklen(op_a.size())];
    let rslt = s.solve((op_c, op_a, op_b, cone, &mut work)).unwrap();
    assert_float_eq!(rslt.0[0], -2., abs_all <= 1e-3);
}
