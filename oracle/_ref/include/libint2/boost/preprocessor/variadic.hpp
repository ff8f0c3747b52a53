# /* **************************************************************************
#  *                                                                          *
#  *     (C) Copyright Edward Diener 2011.                                    *
#  *     (C) Copyright Paul Mensonides 2011.                                  *
#  *     Distributed under the Boost Software License, Version 1.0. (See      *
#  *     accompanying file LICENSE_1_0.txt or copy at                         *
#  *     http://www.boost.org/LICENSE_1_0.txt)                                *
#  *                                                                          *
#  ************************************************************************** */
#
# /* See http://www.boost.org for most recent version. */
#
# ifndef BOOST_PREPROCESSOR_VARIADIC_HPP
# define BOOST_PREPROCESSOR_VARIADIC_HPP
#
# include <libint2/boost/preprocessor/variadic/elem.hpp>
# include <libint2/boost/preprocessor/variadic/size.hpp>
# include <libint2/boost/preprocessor/variadic/to_array.hpp>
# include <libint2/boost/preprocessor/variadic/to_list.hpp>
# include <libint2/boost/preprocessor/variadic/to_seq.hpp>
# include <libint2/boost/preprocessor/variadic/to_tuple.hpp>
#
# endif
