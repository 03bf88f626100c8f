//! Distance metrics: the reference's `Metric` trait and `Euclidean` type, unchanged in shape
//! (reference src/distance.rs:9-55).  The GPU trees accept `Euclidean` only.
use std::ops::AddAssign;

use ndarray::ArrayView1;
use num_traits::{Float, Zero};

pub trait Metric<A> {
    fn distance(&self, _: &ArrayView1<A>, _: &ArrayView1<A>) -> A;
    fn rdistance(&self, _: &ArrayView1<A>, _: &ArrayView1<A>) -> A;
    fn rdistance_to_distance(&self, _: A) -> A;
    fn distance_to_rdistance(&self, _: A) -> A;
}

#[derive(Default, Clone, Debug, Eq, PartialEq)]
pub struct Euclidean {}

unsafe impl Sync for Euclidean {}

impl<A> Metric<A> for Euclidean
where
    A: Float + Zero + AddAssign,
{
    fn distance(&self, x1: &ArrayView1<A>, x2: &ArrayView1<A>) -> A {
        self.rdistance(x1, x2).sqrt()
    }
    fn rdistance(&self, x1: &ArrayView1<A>, x2: &ArrayView1<A>) -> A {
        let mut sum = A::zero();
        for (&a, &b) in x1.iter().zip(x2.iter()) {
            let diff = a - b;
            sum += diff * diff;
        }
        sum
    }
    fn rdistance_to_distance(&self, d: A) -> A {
        d.sqrt()
    }
    fn distance_to_rdistance(&self, d: A) -> A {
        d.powi(2)
    }
}
