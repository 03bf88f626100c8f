//! usage: ref_parity <points.f32.bin> <queries.f32.bin> <n> <nq> <d> <k> <out.bin>
//! Inputs are the little-endian row-major f32 arrays written by
//! `python -m petal_neighbors_b200.synth`-style generators (see tests/); output is
//! nq * k (u64 index, f32 distance) pairs from the unmodified reference.
use std::fs;
use std::io::Write;

use ndarray::{aview1, Array2};
use petal_neighbors::BallTree;

fn read_f32(path: &str, n: usize) -> Vec<f32> {
    let b = fs::read(path).expect("read");
    assert_eq!(b.len(), n * 4);
    b.chunks_exact(4).map(|c| f32::from_le_bytes([c[0], c[1], c[2], c[3]])).collect()
}

fn main() {
    let a: Vec<String> = std::env::args().collect();
    let (n, nq, d, k): (usize, usize, usize, usize) =
        (a[3].parse().unwrap(), a[4].parse().unwrap(), a[5].parse().unwrap(), a[6].parse().unwrap());
    let pts = Array2::from_shape_vec((n, d), read_f32(&a[1], n * d)).unwrap();
    let qs = read_f32(&a[2], nq * d);
    let tree = BallTree::euclidean(pts.view()).expect("non-empty");
    let mut out = fs::File::create(&a[7]).unwrap();
    for q in qs.chunks_exact(d) {
        let (idx, dist) = tree.query(&aview1(q), k);
        for (i, dd) in idx.iter().zip(dist.iter()) {
            out.write_all(&(*i as u64).to_le_bytes()).unwrap();
            out.write_all(&dd.to_le_bytes()).unwrap();
        }
    }
}
